import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, cornell_box
ctx = api.Context(0)
desc = cornell_box(1024, 1024)
rp = RenderParams(spp=256, sampler="sobol", rfilter="box")
p = api.make_params(rp, False, 0, False, 0)
host = np.zeros((1024, 1024, 5), np.float32)
pinned = torch.zeros((1024, 1024, 5), dtype=torch.float32).pin_memory()
for rep in range(3):
    t0 = time.perf_counter(); sc = api.Scene(ctx, desc); t1 = time.perf_counter()
    rc = ctx.L.b2_render(sc.h, C.byref(p), host.ctypes.data_as(C.POINTER(C.c_float))); t2 = time.perf_counter()
    st = sc.stats()
    rc = ctx.L.b2_render(sc.h, C.byref(p), host.ctypes.data_as(C.POINTER(C.c_float))); t3 = time.perf_counter()
    rc = ctx.L.b2_render(sc.h, C.byref(p), C.cast(pinned.data_ptr(), C.POINTER(C.c_float))); t4 = time.perf_counter()
    sc.close(); t5 = time.perf_counter()
    print(f"rep {rep}: scene create+commit {1e3*(t1-t0):.1f} ms | first render wall {1e3*(t2-t1):.1f} (device {st['ms_total']:.1f}) | second render wall {1e3*(t3-t2):.1f} | pinned film {1e3*(t4-t3):.1f} | destroy {1e3*(t5-t4):.1f}")
