import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, smoke_scene
ctx = api.Context(0)
d = smoke_scene(64, 64, res=32)
g = api.Scene(ctx, d)
rp = RenderParams(spp=16, rfilter="box", sampler="independent", integrator="volpath")
for kw in (dict(parity=True, pool_size=1 << 14, flags=8), dict(parity=True, pool_size=1 << 14), dict(parity=False, pool_size=1 << 14), dict(parity=True)):
    try:
        film, st = g.render(rp, **kw)
        print(kw, "ok", api.develop(film).mean(), st["iterations"])
    except Exception as e:
        print(kw, "FAIL", e)
