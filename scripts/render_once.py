#!/usr/bin/env python3
"""One render of a named workload through the C-ABI (profiling target for ncu).  usage: render_once.py <scene> <spp> [res] [flags]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, cornell_box, config3_scene, smoke_scene, stress_scene, textured_scene

name, spp = sys.argv[1], int(sys.argv[2])
res = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = api.Context(0)
kw = dict(sampler="sobol", rfilter=os.environ.get("B2_RFILTER", "box"))
if name == "cornell":
    d = cornell_box(res, res)
elif name == "c3":
    d = config3_scene(res, res)
elif name == "smoke":
    d = smoke_scene(res, res, res=128); kw = dict(sampler="independent", rfilter="gaussian", integrator="volpath")
elif name == "stress":
    d = stress_scene(int(os.environ.get("B2_NINST", "10")), width=res, height=res)
else:
    raise SystemExit("unknown scene")
sc = api.Scene(ctx, d)
_, st = sc.render(RenderParams(spp=spp, **kw), flags=flags)
print(name, spp, res, "Msamples/s", res * res * spp / st["ms_total"] / 1e3, "iterations", st["iterations"])
