"""Host-side BVH build time on the flattened stress scene (no device needed): python scripts/bvh_build_bench.py [instances] [threads]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mitsuba_b200.scene import stress_scene
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 25
threads = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
t = time.time()
d = stress_scene(n_inst, instanced=False)
lo, hi = [], []
for m in d.meshes:
    tri = np.asarray(m.P, np.float32)[np.asarray(m.idx, np.int64).reshape(-1, 3)]
    lo.append(tri.min(axis=1)); hi.append(tri.max(axis=1))
boxes = np.ascontiguousarray(np.concatenate([np.concatenate(lo), np.concatenate(hi)], axis=1), np.float32)
print(f"{len(boxes)} triangles, scene set-up {time.time() - t:.1f} s", flush=True)
L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mitsuba_b200", "libb2mts.so"))
os.environ["B2_COMMIT_TIMING"] = "1"
wide = C.c_uint32()
t = time.time()
n = L.b2_bvh_build_only(boxes.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(len(boxes)), threads, 1, C.byref(wide))
print(f"build: {time.time() - t:.2f} s with {threads} threads, {n} binary nodes, {wide.value} wide nodes")
