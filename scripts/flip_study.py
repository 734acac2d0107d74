#!/usr/bin/env python3
"""GPU probe: which approximation of the throughput build flips paths?  Usage: B2MTS_LIB=<variant .so> python scripts/flip_study.py <tag>.
Prints, per scene, rel-L2 of the throughput build against the parity build, the fraction of pixels with a changed path length and the speed."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, cornell_box, material_ball, config3_scene
import test_gpu_parity as T


def rel_l2(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).sum() / (b.astype(np.float64) ** 2).sum()))


ctx = api.Context(0)
out = {"tag": sys.argv[1] if len(sys.argv) > 1 else "?", "lib": os.environ.get("B2MTS_LIB", "default")}
for name, d, spp in (("roughdielectric_ggx_128", material_ball(T.MATERIALS["roughdielectric_ggx"], 128, 128, 48, 96), 256),
                     ("c3_256", config3_scene(256, 256, 100, 200), 256)):
    g = api.Scene(ctx, d)
    rp = RenderParams(spp=spp, sampler="sobol", rfilter="box")
    fp, sp = g.render(rp, parity=True, flags=32); pp = g.pixel_stats()
    ff, sf = g.render(rp, parity=False, flags=32); pf = g.pixel_stats()
    W, H = g.W, g.H
    out[name] = {"fast_vs_parity": rel_l2(api.develop(ff), api.develop(fp)), "changed_paths_lower_bound": float((pp != pf).sum() / (W * H * spp)),
                 "ms_fast": sf["ms_total"], "ms_parity": sp["ms_total"]}
    g.close()
for name, d, spp in (("c2", cornell_box(1024, 1024), 256), ("c3", config3_scene(1024, 1024), 128)):
    g = api.Scene(ctx, d)
    rp = RenderParams(spp=spp, sampler="sobol", rfilter="box")
    g.render(rp, parity=False)
    _, sf = g.render(rp, parity=False)
    out[name + "_msamples_s"] = 1024 * 1024 * spp / sf["ms_total"] / 1e3
    g.close()
print(json.dumps(out))
