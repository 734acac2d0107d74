"""First GPU contact: ray parity, render parity and a rough throughput number on the Cornell scene."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mitsuba_b200.scene import cornell_box, RenderParams
from mitsuba_b200 import api
from oracle import oracle_api as O

def rel_l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))

ctx = api.Context(0)
d = cornell_box(128, 128)
sc = api.Scene(ctx, d)
osc = O.OracleScene(d)
print("stats", sc.stats())
# camera parity
rng = np.random.default_rng(1)
pos = rng.uniform(0, 128, (4096, 2)).astype(np.float32)
rg = sc.camera_rays(pos); ro = osc.camera_rays(pos)
print("camera rays max abs diff", np.abs(rg - ro)[np.isfinite(ro)].max())
# trace parity: camera rays + random interior rays
t0, u0, v0, p0 = osc.trace(ro, 0)
t1, u1, v1, p1 = sc.trace(ro, 0, parity=True)
print("closest: prim mismatch", int((p0 != p1).sum()), "of", len(p0), "max |dt|", float(np.abs(t0 - t1)[p0 == p1][np.isfinite(t0[p0 == p1])].max()))
o = rng.uniform(10, 540, (20000, 3)).astype(np.float32)
dd = rng.normal(size=(20000, 3)).astype(np.float32); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
rays = np.concatenate([o, np.full((20000, 1), 1e-4, np.float32), dd, np.full((20000, 1), np.inf, np.float32)], 1).astype(np.float32)
t0, u0, v0, p0 = osc.trace(rays, 0); t1, u1, v1, p1 = sc.trace(rays, 0, parity=True)
same = p0 == p1
print("random closest: prim mismatch", int((~same).sum()), "max|dt|", float(np.abs(t0 - t1)[same & np.isfinite(t0)].max()), "max|du|", float(np.abs(u0 - u1)[same].max()))
rays[:, 7] = rng.uniform(50, 600, 20000)
_, _, _, q0 = osc.trace(rays, 1); _, _, _, q1 = sc.trace(rays, 1, parity=True)
print("occlusion mismatch", int((q0 != q1).sum()))
# sampler stream
for (px, py, s) in [(0, 0, 0), (5, 77, 3), (127, 127, 15)]:
    a = sc.sampler_stream("sobol", 0, 16, px, py, s, 12); b = osc.sampler_stream("sobol", 0, 16, px, py, s, 12)
    print("sobol stream equal", np.array_equal(a, b))
# render parity
for filt in ("box", "gaussian"):
    rp = RenderParams(spp=16, sampler="sobol", rfilter=filt)
    fo, so = osc.render(rp)
    for parity in (True, False):
        t = time.time()
        fg, sg = sc.render(rp, parity=parity)
        dt = time.time() - t
        rgb_g, rgb_o = api.develop(fg), O.develop(fo)
        print(filt, "parity" if parity else "fast", "relL2 %.3e" % rel_l2(rgb_g, rgb_o), "weight diff", float(np.abs(fg[..., 4] - fo[..., 4]).max()),
              "wall %.3fs" % dt, {k: sg[k] for k in ("samples", "rays", "shadow_rays", "path_length_sum", "iterations", "ms_total")})
    print("  oracle", {k: so[k] for k in ("samples", "rays", "shadowRays", "pathLengthSum")})
# throughput
d2 = cornell_box(512, 512)
sc2 = api.Scene(ctx, d2)
for pool in (1 << 18, 1 << 20, 1 << 22):
    rp = RenderParams(spp=64, sampler="sobol", rfilter="box")
    sc2.render(rp, parity=False, pool_size=pool)
    fg, sg = sc2.render(rp, parity=False, pool_size=pool)
    print("pool", pool, "ms", sg["ms_total"], "Msamples/s %.1f" % (512 * 512 * 64 / sg["ms_total"] / 1e3), "iters", sg["iterations"])
import cv2
rgb = api.develop(fg)
cv2.imwrite("gpurun_out/cbox_gpu.png", (np.clip(rgb, 0, 1) ** (1 / 2.2) * 255)[..., ::-1].astype(np.uint8))
