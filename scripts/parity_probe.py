#!/usr/bin/env python3
"""GPU probe (not a test): how far is the throughput build (FMA, --use_fast_math) from the parity build and from the oracle, as a function
of the sample count, and what fraction of pixels holds a path whose length differs (b2_get_pixel_stats)?  Prints one JSON object."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, cornell_box, material_ball, config3_scene
from oracle import oracle_api as O
import test_gpu_parity as T


def rel_l2(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).sum() / (b.astype(np.float64) ** 2).sum()))


ctx = api.Context(0)
out = {"small": {}, "large": {}}
scenes = {"cornell": cornell_box(64, 64), "c3_two_balls": config3_scene(64, 64, 48, 96)}
for k in ("roughconductor_ggx", "roughdielectric_ggx", "roughdielectric_beckmann", "coating_roughconductor"):
    scenes[k] = material_ball(T.MATERIALS[k], 64, 64, 48, 96)
for name, d in scenes.items():
    g = api.Scene(ctx, d)
    o = O.OracleScene(d, sample_to_camera=g.sample_to_camera())
    rows = {}
    for spp in (32, 256, 2048):
        rp = RenderParams(spp=spp, sampler="sobol", rfilter="gaussian")
        t0 = time.time(); fo, so = o.render(rp); t_or = time.time() - t0
        fp, sp = g.render(rp, parity=True, flags=32); pp = g.pixel_stats()
        ff, sf = g.render(rp, parity=False, flags=32); pf = g.pixel_stats()
        a, b, c = api.develop(fp), api.develop(ff), O.develop(fo)
        rows[spp] = {"parity_vs_oracle": rel_l2(a, c), "fast_vs_oracle": rel_l2(b, c), "fast_vs_parity": rel_l2(b, a),
                     "pixels_with_changed_paths": float((pp != pf).mean()), "changed_paths_lower_bound": float((pp != pf).sum() / (64 * 64 * spp)),
                     "oracle_s": t_or, "len_parity": sp["path_length_sum"] / sp["samples"], "len_fast": sf["path_length_sum"] / sf["samples"],
                     "len_oracle": so["pathLengthSum"] / so["samples"]}
    out["small"][name] = rows
    g.close()
# the benchmarked sizes, device only: throughput build against parity build
for name, d, spp in (("c2_cornell_1024x1024", cornell_box(1024, 1024), 1024), ("c3_two_balls_1024x1024", config3_scene(1024, 1024), 512)):
    g = api.Scene(ctx, d)
    rp = RenderParams(spp=spp, sampler="sobol", rfilter="box")
    fp, sp = g.render(rp, parity=True)
    ff, sf = g.render(rp, parity=False)
    a, b = api.develop(fp), api.develop(ff)
    out["large"][name] = {"spp": spp, "fast_vs_parity": rel_l2(b, a), "ms_parity": sp["ms_total"], "ms_fast": sf["ms_total"],
                          "msamples_s_fast": 1024 * 1024 * spp / sf["ms_total"] / 1e3, "len_parity": sp["path_length_sum"] / sp["samples"], "len_fast": sf["path_length_sum"] / sf["samples"]}
    g.close()
print(json.dumps(out, indent=1))
