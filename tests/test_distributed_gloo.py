"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank renders its sample-index shard (here with the
oracle standing in for the device renderer), the films are SUM-reduced to rank 0 and must equal the
single-process film (weight channel exactly, RGB to float summation order)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mitsuba_b200.distributed import render_sharded, shard_range, torch_reduce_sum
from mitsuba_b200.scene import RenderParams, cornell_box, smoke_scene


def test_shard_ranges_tile_the_sample_range():
    for spp in (2, 7, 64, 1024):
        for world in (1, 2, 4, 8):
            if spp < world:
                with pytest.raises(ValueError):
                    shard_range(spp, 0, world)
                continue
            rs = [shard_range(spp, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == spp
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:])) and all(b > a for a, b in rs)
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    assert shard_range(64, 1, 2, lo=16, hi=48) == (32, 48)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _case(kind):
    if kind == "volpath":   # the medium does not change the partitioning (independent (pixel, sample) units)
        return smoke_scene(24, 24, res=12), RenderParams(spp=6, rfilter="gaussian", sampler="independent", integrator="volpath")
    if kind == "envmap":    # an environment map is scene data replicated on every rank, like the geometry
        from mitsuba_b200.scene import envmap_scene
        return envmap_scene(24, 24, map_width=32, n_theta=8, n_phi=16), RenderParams(spp=6, rfilter="gaussian")
    return cornell_box(32, 32), RenderParams(spp=6, rfilter="gaussian")


def _worker(rank, world, port, out_path, kind="path"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle_api as O
    desc, rp = _case(kind)
    sc = O.OracleScene(desc)

    def render_fn(shard):
        film, _ = sc.render(shard, threads=1)
        return torch.from_numpy(film)

    film = render_sharded(render_fn, rp, rank, world, lambda f: torch_reduce_sum(f, 0))
    if rank == 0:
        np.save(out_path, film.numpy())
    else:
        assert film is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["path", "volpath", "envmap"])
def test_two_rank_gloo_reduce_equals_single_process(tmp_path, kind):
    out = str(tmp_path / "film.npy")
    port = 29500 + (os.getpid() + (7 if kind == "volpath" else 0)) % 2000
    mp.spawn(_worker, args=(2, port, out, kind), nprocs=2, join=True)
    film2 = np.load(out)
    from oracle import oracle_api as O
    desc, rp = _case(kind)
    full, _ = O.OracleScene(desc).render(rp, threads=1)
    assert np.allclose(film2, full, rtol=1e-5, atol=1e-6)
    assert np.allclose(film2[..., 4], full[..., 4], rtol=1e-6)
