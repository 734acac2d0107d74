#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json): Msamples/s of the Cornell-box-class scene,
1024 x 1024 @ 1024 spp per GPU, `path` integrator, sobol sampler, box reconstruction filter.

    python bench.py --gpus N --steps K --warmup W            # this implementation (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference on the host cores

One "step" = one complete render of the workload (all pixels x all samples of this rank's shard) + the film reduce.
N > 1: weak scaling -- every rank renders its own 1024 sample indices of every pixel (rank r: [r*1024, (r+1)*1024) of a
1024*N-spp image), one torch.distributed reduce(SUM) of the (H, W, 5) film over NCCL at the end of every step.
`value` = samples rendered by all ranks / max-over-ranks device time.  `e2e` = the same metric through the C-ABI with HOST
buffers: scene description -> b2_scene_commit (BVH build + H2D upload) -> b2_render into a host film (D2H) inside the timed
region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(scene="cornell_box S1 (32 triangles, diffuse + area light)", width=1024, height=1024, spp_per_gpu=1024,
                integrator="path maxDepth=-1 rrDepth=5", sampler="sobol scramble=0", rfilter="box")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


REF_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libpathref.so")
_REF = {}


def _ref_worker_init(width, height, spp, scene="cornell"):
    """One scene per worker process: the reference's own Scene / ShapeKDTree / MIPathTracer or VolumetricPathTracer (oracle/path_ref_shim.cpp)."""
    import ctypes as C
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import ref_pins
    from mitsuba_b200.scene import RenderParams, cornell_box, smoke_scene
    lib = C.CDLL(REF_LIB)
    if scene == "smoke":  # BASELINE configs[3]: 128^3 heterogeneous medium, volpath, this repository's counter stream as the sampler
        d, rp = smoke_scene(width, height, res=128), RenderParams(spp=spp, sampler="independent", rfilter="gaussian", integrator="volpath")
    else:
        d, rp = cornell_box(width, height), RenderParams(spp=spp, sampler="sobol", rfilter="box")
    _REF["lib"], _REF["handle"], _REF["shape"] = lib, ref_pins.reference_scene(lib, d, rp), (height, width, 5)


def reference_cpu_rate(scene, width, height, spp, steps=1, warmup=1, cores=0):
    """Msamples/s of the reference's own code (oracle/_ref/libpathref.so) on `cores` processes, or None where the library is absent."""
    if not os.path.exists(REF_LIB) or os.environ.get("B2_BENCH_ORACLE_PORT"):
        return None
    import multiprocessing as mp
    cores = cores or os.cpu_count()
    with mp.get_context("fork").Pool(cores, initializer=_ref_worker_init, initargs=(width, height, spp, scene)) as pool:
        shares = [(i, cores) for i in range(cores)]
        for _ in range(warmup):
            pool.map(_ref_worker_render, shares)
        t0 = time.time()
        for _ in range(steps):
            w = pool.map(_ref_worker_render, shares)
        dt = (time.time() - t0) / max(steps, 1)
    n = width * height * spp
    if abs(sum(w) - n) > 2e-2 * n:  # film weights (gaussian: the border pixels lose a little)
        raise RuntimeError("the shares of the reference render do not add up to the whole image")
    return dict(value=n / dt / 1e6, unit="Msamples/s", cores=cores, kind="reference", ms_per_step=dt * 1e3)


def _ref_worker_render(args):
    import ctypes as C
    first, step = args
    film = np.zeros(_REF["shape"], np.float32)
    _REF["lib"].pathref_render_blocks(_REF["handle"], first, step, film.ctypes.data_as(C.POINTER(C.c_float)))
    return float(film[..., 4].sum())


def cpu_reference_run(steps, warmup, sample_spp=None, threads=0):
    """The reference on the host cores, on a bounded sample of the SAME workload (`sample_spp` samples of every pixel of the
    1024 x 1024 image).  kind "reference": the reference's own sources (path.cpp, scene.cpp, skdtree.cpp, the plugins ...) compiled into
    oracle/_ref/libpathref.so, one process per core, each rendering every cores-th 32 x 32 block with SamplingIntegrator::renderBlock
    (the reference's scheduler is not part of that build).  Falls back to the oracle port (kind "port") where that library is absent."""
    sample_spp = sample_spp or 16
    r = reference_cpu_rate("cornell", WORKLOAD["width"], WORKLOAD["height"], sample_spp, steps, max(warmup, 1), threads)
    if r is not None:
        n = WORKLOAD["width"] * WORKLOAD["height"] * sample_spp
        r.update(sample=f"{sample_spp} spp (Sobol', box filter) of every pixel of the 1024x1024 Cornell workload ({n / 1e6:.1f} Msamples per step), "
                        "reference sources compiled into oracle/_ref/libpathref.so, one process per core", mean_path_length=None)
        return r
    return cpu_port_run(steps, warmup, sample_spp, threads)


def cpu_port_run(steps, warmup, sample_spp=None, threads=0):
    """The CPU restatement of the reference (oracle, kind "port") on the host cores: a bounded sample of the SAME workload
    (the first `sample_spp` sample indices of every pixel of the 1024 x 1024 image)."""
    from mitsuba_b200.scene import RenderParams, cornell_box
    from oracle import oracle_api as O
    cores = threads or O.hardware_threads()
    d = cornell_box(WORKLOAD["width"], WORKLOAD["height"])
    sc = O.OracleScene(d)
    if sample_spp is None:
        # calibrate so one step is ~4-10 s of wall time
        rp = RenderParams(spp=WORKLOAD["spp_per_gpu"], sampler="sobol", rfilter="box", sample_lo=0, sample_hi=1)
        t = time.time(); sc.render(rp, threads=cores); dt = time.time() - t
        sample_spp = int(min(16, max(1, round(5.0 / max(dt, 1e-3)))))
    rp = RenderParams(spp=WORKLOAD["spp_per_gpu"], sampler="sobol", rfilter="box", sample_lo=0, sample_hi=sample_spp)
    for _ in range(warmup):
        sc.render(rp, threads=cores)
    t0 = time.time()
    st = None
    for _ in range(steps):
        _, st = sc.render(rp, threads=cores)
    dt = (time.time() - t0) / max(steps, 1)
    n = WORKLOAD["width"] * WORKLOAD["height"] * sample_spp
    return dict(value=n / dt / 1e6, unit="Msamples/s", cores=cores, kind="port",
                sample=f"sample indices [0,{sample_spp}) of every pixel of the 1024x1024 @1024spp workload ({n / 1e6:.1f} Msamples per step)",
                ms_per_step=dt * 1e3, mean_path_length=st["pathLengthSum"] / st["samples"])


def traversal_metric(ctx, hbm_gbs, n_inst=10, n_rays=1 << 22):
    """Second half of BASELINE's metric ("traversal HBM GB/s vs peak"): k_trace_rays (the traversal loop of k_extend) on an
    S3-class scene that is NOT cache resident -- 1 M triangles (10 x 100 k instanced, flattened) -- with incoherent rays from
    the bounding sphere (kdbench-style origins, src/utils/kdbench.cpp:222-229) aimed at random mesh vertices.  Algorithmic bytes per ray = 48 (ray in, hit out) + 64 B per
    node visit + 48 B per triangle test, visits/tests counted by the kernel itself."""
    import torch
    from mitsuba_b200 import api
    from mitsuba_b200.scene import stress_scene
    d = stress_scene(n_inst, width=64, height=64)
    sc = api.Scene(ctx, d)
    P = np.concatenate([m.P for m in d.meshes]); lo, hi = P.min(0), P.max(0)
    g = torch.Generator(device="cuda").manual_seed(0)

    def sph():
        v = torch.randn((n_rays, 3), device="cuda", generator=g)
        return v / v.norm(dim=1, keepdim=True)
    c = torch.tensor((lo + hi) / 2, device="cuda", dtype=torch.float32); r = float(np.linalg.norm(hi - lo) / 2)
    # origin: uniform on the bounding sphere (kdbench); target: a random mesh vertex of the instanced geometry, so that every ray
    # descends to the leaves instead of ending on the two ground triangles
    Pg = torch.tensor(np.concatenate([m.P for m in d.meshes[:-2]]), device="cuda")
    a = c + r * sph()
    b = Pg[torch.randint(0, len(Pg), (n_rays,), device="cuda", generator=g)]
    dd = b - a; L = dd.norm(dim=1, keepdim=True)
    rays = torch.cat([a, torch.zeros((n_rays, 1), device="cuda"), dd / L, 2 * L], 1).contiguous().float()
    out = torch.zeros((n_rays, 4), device="cuda")
    sc.trace_device(rays, out, n_rays, mode=2)
    st = sc.stats()
    nv, pt = st["node_visits"] / n_rays, st["prim_tests"] / n_rays
    for _ in range(3):
        sc.trace_device(rays, out, n_rays, mode=0)
    ms = min(sc.trace_device(rays, out, n_rays, mode=0) for _ in range(5))
    alg = (48 + 64 * nv + 48 * pt) * n_rays
    res = {"scene": f"stress {n_inst}x100k = {d.n_triangles()} triangles, BVH2 {st['n_bvh_nodes']} nodes", "rays": n_rays, "kernel": "k_trace_rays (closest hit)",
           "mrays_s": n_rays / ms / 1e3, "ms": ms, "node_visits_per_ray": nv, "tri_tests_per_ray": pt, "achieved": alg / ms / 1e6, "unit": "GB/s",
           "peak": hbm_gbs, "frac": alg / ms / 1e6 / hbm_gbs,
           "note": "algorithmic bytes; the top of the tree is served from shared memory / L1 / L2, so DRAM traffic is far lower (profiles/)"}
    sc.close()
    return res


def volpath_metric(ctx, with_cpu=True, cpu_ref=None):
    """BASELINE configs[3], reported next to the headline: the S4 smoke scene -- a 128^3 density grid in the unit cube,
    `heterogeneous` Woodcock medium, isotropic phase, `volpath`, 512x512 @ 256 spp -- on this rank's GPU, plus the oracle's rate for
    the same scene on the host cores (bounded sample)."""
    from mitsuba_b200 import api
    from mitsuba_b200.scene import RenderParams, smoke_scene
    d = smoke_scene(512, 512, res=128)
    sc = api.Scene(ctx, d)
    rp = RenderParams(spp=256, rfilter="gaussian", sampler="independent", integrator="volpath")
    sc.render(RenderParams(spp=16, rfilter="gaussian", sampler="independent", integrator="volpath"))
    best = None
    for _ in range(2):
        _, st = sc.render(rp, flags=4)
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    n = 512 * 512 * 256
    res = {"workload": "S4 smoke: 128^3 gridvolume, heterogeneous (woodcock), isotropic, volpath, independent sampler, gaussian filter, 512x512 @ 256 spp, 1 GPU",
           "value": n / best["ms_total"] / 1e3, "unit": "Msamples/s", "ms": best["ms_total"], "kernel": "k_volstep", "kernel_ms": best["ms_shade"],
           "mean_path_length": best["path_length_sum"] / best["samples"], "rays_per_sample": (best["rays"] + best["shadow_rays"]) / best["samples"]}
    sc.close()
    if cpu_ref is not None:
        res["cpu_baseline"] = {**{k: cpu_ref[k] for k in ("value", "unit", "cores", "kind")},
                               "sample": "8 spp of every pixel of the same scene; the reference's volpath.cpp / heterogeneous.cpp / gridvolume.cpp compiled into oracle/_ref/libpathref.so, one process per core"}
    elif with_cpu:
        from oracle import oracle_api as O
        o = O.OracleScene(smoke_scene(512, 512, res=128))
        rp2 = RenderParams(spp=8, rfilter="gaussian", sampler="independent", integrator="volpath")
        t0 = time.time(); _, so = o.render(rp2); dt = time.time() - t0
        res["cpu_baseline"] = {"value": so["samples"] / dt / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "sample indices [0,8) of every pixel of the same scene"}
    return res


def textured_metric(ctx, with_cpu=True):
    """SURVEY.md 8f-4, reported next to the headline: the textured material-ball scene (three `bitmap` textures of 1024^2 texels, EWA
    filtering through ray differentials on camera hits, bilinear on the bounces), `path`, 1024x1024 @ 64 spp on this rank's GPU, plus
    the oracle's rate for the same scene on the host cores (bounded sample)."""
    from mitsuba_b200 import api
    from mitsuba_b200.scene import RenderParams, textured_scene
    mk = lambda: textured_scene(1024, 1024, filter_type="ewa", tex_res=1024, n_theta=200, n_phi=200)
    sc = api.Scene(ctx, mk())
    rp = RenderParams(spp=64, rfilter="gaussian", sampler="sobol")
    sc.render(RenderParams(spp=4, rfilter="gaussian", sampler="sobol"))
    best = None
    for _ in range(2):
        _, st = sc.render(rp, flags=4)
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    n = 1024 * 1024 * 64
    res = {"workload": "S2 material ball, 3 bitmap textures (1024^2, ewa, maxAnisotropy 20), ~80k triangles, path, sobol, gaussian filter, 1024x1024 @ 64 spp, 1 GPU",
           "value": n / best["ms_total"] / 1e3, "unit": "Msamples/s", "ms": best["ms_total"], "kernel": "k_shade<-1, false, TEX>", "kernel_ms": best["ms_shade"],
           "mean_path_length": best["path_length_sum"] / best["samples"]}
    sc.close()
    if with_cpu:
        from oracle import oracle_api as O
        o = O.OracleScene(mk())
        t0 = time.time(); _, so = o.render(RenderParams(spp=4, rfilter="gaussian", sampler="sobol")); dt = time.time() - t0
        res["cpu_baseline"] = {"value": so["samples"] / dt / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "sample indices [0,4) of every pixel of the same scene"}
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": "Msamples/sec Cornell box 1024spp", "value": r["value"], "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Cornell box (S1) 1024x1024, path/sobol/box, bounded sample on the host cores", **WORKLOAD},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": ("the reference's own sources (MIPathTracer::Li, renderBlock, Scene, ShapeKDTree, plugins) compiled into oracle/_ref/libpathref.so; "
                     "the Mitsuba binary itself (SCons, Boost, Xerces, OpenEXR) cannot be built offline (DESIGN.md)") if r["kind"] == "reference" else
                    "Mitsuba-0.6-equivalent CPU restatement (oracle/), not the Mitsuba binary: the reference cannot be built offline (DESIGN.md)"}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp_per_gpu"], help="samples per pixel per GPU (headline: 1024)")
    ap.add_argument("--res", type=int, default=WORKLOAD["width"])
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--parity", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-volpath", action="store_true", help="skip the volpath (BASELINE configs[3]) side measurement")
    ap.add_argument("--no-traversal", action="store_true", help="skip the isolated BVH-traversal measurement (S3-class scene)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # CPU arm first: the reference renderer forks one process per core, which must happen before this process owns a CUDA context
    cpu_base = None
    if not args.no_cpu_baseline and args.gpus == 1 and int(os.environ.get("RANK", "0")) == 0:
        try:
            cpu_base = cpu_reference_run(1, 1)
        except Exception as e:
            print(f"[bench] reference CPU arm failed ({e}); falling back to the oracle port", file=sys.stderr)
            cpu_base = cpu_port_run(1, 0)
    vol_cpu = None
    if not args.no_cpu_baseline and not args.no_volpath and args.gpus == 1 and int(os.environ.get("RANK", "0")) == 0:
        try:
            vol_cpu = reference_cpu_rate("smoke", 512, 512, 8)
        except Exception as e:
            print(f"[bench] reference CPU arm (volpath) failed ({e}); the oracle port is timed instead", file=sys.stderr)

    import torch
    import torch.distributed as dist
    from mitsuba_b200 import api
    from mitsuba_b200.distributed import shard_range
    from mitsuba_b200.scene import RenderParams, cornell_box

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this implementation has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    W = H = args.res
    desc = cornell_box(W, H)
    scene = api.Scene(ctx, desc)
    total_spp = args.spp * world
    lo, hi = shard_range(total_spp, rank, world)
    rp = RenderParams(spp=total_spp, sampler="sobol", rfilter="box", sample_lo=lo, sample_hi=hi)
    film = torch.zeros((H, W, 5), dtype=torch.float32, device=f"cuda:{local}")

    def step(flags=4):
        scene.render(rp, parity=bool(args.parity), pool_size=args.pool, flags=flags, film=film)
        if world > 1:
            dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)   # the one collective of the path: film merge over NVLink
        return scene.stats()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    agg = dict(ms_generate=0.0, ms_extend=0.0, ms_shade=0.0, ms_occluded=0.0, n_generate=0, n_extend=0, n_shade=0, n_occluded=0,
               launches=0, rays=0, shadow_rays=0, path_length_sum=0, samples=0, ms_render=0.0, unoccluded=0)
    t_wall = time.time()
    for _ in range(args.steps):
        st = step()
        for k in ("ms_generate", "ms_extend", "ms_shade", "ms_occluded", "n_generate", "n_extend", "n_shade", "n_occluded", "rays", "shadow_rays",
                  "path_length_sum", "samples"):
            agg[k] += st[k]
        agg["unoccluded"] += st["unoccluded_shadow_rays"]
        agg["launches"] += st["kernel_launches"] + (1 if world > 1 else 0)
        agg["ms_render"] += st["ms_total"]
    ev1.record()
    sync()
    wall = time.time() - t_wall
    # b2_render times itself with CUDA events on ITS stream (torch's events only see torch's stream); each b2_render call
    # synchronises its stream before returning, so the per-step device time = render ms (+ reduce, measured by torch events)
    ms_torch = ev0.elapsed_time(ev1)
    ms_total = max(ms_torch, agg["ms_render"])
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clock_info = clocks.stop() if clocks else None
    samples_per_step = W * H * args.spp * world
    value = samples_per_step * args.steps / (ms_total / 1e3) / 1e6
    pool = scene.stats()["pool_size"]

    # ---- e2e: through the C-ABI with host buffers (scene commit + render into a host film), same metric ----
    e2e = None
    if rank == 0 or world > 1:
        host_film = torch.zeros((H, W, 5), dtype=torch.float32).pin_memory()   # the step's result is read back into pinned host memory
        import ctypes as C
        p = api.make_params(rp, bool(args.parity), args.pool, False, 0)

        def e2e_step():
            sc2 = api.Scene(ctx, desc)                                  # b2_scene_create .. b2_scene_commit (H2D upload)
            rc = ctx.L.b2_render(sc2.h, C.byref(p), C.cast(host_film.data_ptr(), C.POINTER(C.c_float)))  # D2H film inside
            if rc:
                raise RuntimeError(ctx.err())
            up = sc2.stats()["bytes_uploaded"]
            sc2.close()
            return up
        e2e_step()
        sync()
        t0 = time.time()
        n_e2e = max(1, min(args.steps, 2))
        up = 0
        for _ in range(n_e2e):
            up = e2e_step()
        sync()
        dt = time.time() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": samples_per_step * n_e2e / float(tt.item()) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(up) + 64,
               "d2h_bytes_per_step": int(H * W * 5 * 4), "steps": n_e2e}

    # cross-check of the in-kernel %globaltimer stamps: one extra, untimed step with plain launches bracketed by CUDA events
    ev_check = None
    if rank == 0:
        # (render only -- no collective here: the other ranks do not take part)
        _, stc = scene.render(rp, parity=bool(args.parity), pool_size=args.pool, flags=4 | 8, film=film)
        ev_check = {k: stc["ms_" + k] / max(1, stc["n_" + k]) for k in ("generate", "extend", "shade", "occluded")}
    if rank == 0:
        peaks, which = measured_peaks()
        # dominant kernel by summed device time inside the timed region
        shares = {k: agg["ms_" + k] for k in ("generate", "extend", "shade", "occluded")}
        dom = max(shares, key=shares.get)
        n_dom = max(1, agg["n_" + dom])
        avg_ms = shares[dom] / n_dom
        # algorithmic bytes per launch (DESIGN.md "roofline model"): per live path / shadow ray and launch
        # algorithmic bytes per item of each stage ("roofline model"; 32-byte pool records)
        rays, shadow, smp = max(1, agg["rays"]), max(1, agg["shadow_rays"]), max(1, agg["samples"])
        per_item = {"extend": 4 + 32 + 16,
                    "occluded": 16 + 16 + 16 + 32 * agg["unoccluded"] / shadow,
                    "shade": (4 + 16 + 16 + 16 + 8) + (32 + 16 + 4) + 32 * shadow / rays + 4 * smp / rays,
                    "generate": (4 + 4 + 4 + 16 + 8) + 2 * 20 + (32 + 32 + 8 + 8 + 4 + 4)}[dom]
        items = {"extend": rays, "occluded": shadow, "shade": rays, "generate": smp}[dom] / n_dom
        achieved = per_item * items / (avg_ms / 1e3) / 1e9 if avg_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_" + dom)
            except Exception:
                pass
        line = {
            "metric": "Msamples/sec Cornell box 1024spp", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Cornell box (S1) {W}x{H} @ {args.spp} spp per GPU, path/sobol/box (BASELINE configs[1])", **WORKLOAD,
                       "spp_per_gpu": args.spp, "width": W, "height": H, "parallelism": f"sample-index sharding x{world}, 1 film reduce",
                       "pool_size": int(pool), "l2": "path pool + film (%.0f MB) exceed the 126 MB L2; every iteration re-streams them" % ((pool * (64 + 16 + 8 + 8 + 8 + 32 + 8) + W * H * 20) / 1e6),
                       "fp": "parity (-fmad=false)" if args.parity else "fast (FMA contraction)"},
            "e2e": e2e,
            "gpu_launches": int(agg["launches"]),
            "clocks": clock_info,
            "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                         "frac": achieved / peaks.get("hbm_gbs", 1.0), "traffic": traffic, "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)",
                         "avg_launch_ms": avg_ms, "avg_launch_ms_cuda_events": ev_check["" + dom] if ev_check else None,
                         "timer": "%globaltimer stamps inside the kernels (max CTA end - min CTA start per launch), timed region runs as one CUDA graph per iteration; cross-checked by an untimed step with CUDA events around plain launches",
                         "share_of_step": shares[dom] / max(1e-9, sum(shares.values())),
                         "kernel_ms": shares,
                         "note": "Cornell scene (3.5 KB) is shared-memory resident: traversal is issue/latency bound, HBM traffic is queue traffic only"},
            "stats": {"mean_path_length": agg["path_length_sum"] / max(1, agg["samples"]), "rays_per_sample": agg["rays"] / max(1, agg["samples"]),
                      "shadow_rays_per_sample": agg["shadow_rays"] / max(1, agg["samples"]), "wall_s": wall},
        }
        if not args.no_traversal:
            line["traversal"] = traversal_metric(ctx, peaks.get("hbm_gbs", 6650.0))
        if not args.no_volpath:
            line["volpath"] = volpath_metric(ctx, with_cpu=not args.no_cpu_baseline, cpu_ref=vol_cpu)
            try:
                line["textured"] = textured_metric(ctx, with_cpu=not args.no_cpu_baseline)
            except Exception as e:  # a side measurement must not take the headline line down with it
                line["textured"] = {"error": str(e)}
        if cpu_base is not None:
            line["cpu_baseline"] = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
