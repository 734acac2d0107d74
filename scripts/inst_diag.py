import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from mitsuba_b200 import api
from mitsuba_b200.scene import *
from oracle import oracle_api as O
import test_gpu_parity as T
ctx = api.Context(0)
def rel(a, b): return float(np.sqrt(((a.astype(np.float64) - b) ** 2).sum() / (b.astype(np.float64) ** 2).sum()))
rp = RenderParams(spp=64, sampler="sobol", rfilter="box")
for tag in ("asis", "diffuse", "nouv", "nogroup1", "noshear"):
    d = T._instanced_scene()
    if tag == "diffuse": d.meshes[-2].bsdf = Bsdf("diffuse")
    if tag == "nouv": d.meshes[-2].UV = None
    if tag == "nogroup1": d.instances = [i for i in d.instances if i.group == 0]
    if tag == "noshear":
        for i in d.instances:
            if i.group == 1: i.to_world = np.eye(4, dtype=np.float32) * 1.0; i.to_world[:3, 3] = (0.5, 2.6, -1.0)
    g = api.Scene(ctx, d); o = O.OracleScene(d, sample_to_camera=g.sample_to_camera())
    fo, _ = o.render(rp); fp, _ = g.render(rp, parity=True); ff, _ = g.render(rp, parity=False)
    ro, rpp, rf = O.develop(fo), api.develop(fp), api.develop(ff)
    diff = np.abs(rf - ro).sum(2); k = np.unravel_index(np.argmax(diff), diff.shape)
    print(tag, "parity", rel(rpp, ro), "fast", rel(rf, ro), "max diff at", k, float(diff.max()), "n>0.05:", int((diff > 0.05).sum()))
