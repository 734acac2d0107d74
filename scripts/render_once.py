#!/usr/bin/env python3
"""One render of a named workload through the C-ABI (profiling target for ncu).  usage: render_once.py <scene> <spp> [res] [flags]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, cornell_box, config3_scene, envmap_scene, smoke_scene, stress_scene, textured_scene

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
elif name == "envmap":      # config-3 balls under the synthetic sky only
    d = envmap_scene(res, res)
elif name == "envconst":    # the same geometry under a constant environment (what the map costs on top)
    d = envmap_scene(res, res); d.envmap = None; d.env_radiance = (0.5, 0.6, 0.8)
elif name == "f3mix":       # config 3 with a twosided ground and a plastic third ball: BSDFs with and without a specialised shading instance
    from mitsuba_b200.scene import Bsdf, Mesh, uv_sphere
    d = config3_scene(res, res)
    d.meshes[0].bsdf = Bsdf("twosided", nested=Bsdf("diffuse", reflectance=(0.5, 0.5, 0.5)))
    P, N, UV, I = uv_sphere((0.0, 0.6, -1.6), 0.6, 120, 120, smooth=True)
    d.meshes.insert(3, Mesh(P, I, N=N, bsdf=Bsdf("plastic", int_ior=1.49, diffuse_reflectance=(0.1, 0.27, 0.36)), name="ball_plastic"))
elif name == "textured":
    d = textured_scene(res, res, filter_type="ewa", tex_res=1024, n_theta=200, n_phi=200)
else:
    raise SystemExit("unknown scene")
sc = api.Scene(ctx, d)
_, st = sc.render(RenderParams(spp=spp, **kw), flags=flags)
print(name, spp, res, "flags", flags, "Msamples/s", res * res * spp / st["ms_total"] / 1e3, "iterations", st["iterations"], "ms_shade", st["ms_shade"], "ms_extend", st["ms_extend"], "ms_occluded", st["ms_occluded"], "ms_generate", st["ms_generate"])
