#!/usr/bin/env python3
"""GPU probe: WHERE do paths of the throughput build leave the parity build's path?  Per-sample event traces (b2_get_path_traces) of both
builds on the config-3 scene; for every sample whose traces differ, the first differing bounce and which event field differs."""
import collections, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, config3_scene, material_ball
import test_gpu_parity as T

ctx = api.Context(0)
res, spp = 256, 256
out = {}
import itertools
for (name, d), (mode, mflags) in itertools.product((("c3", config3_scene(res, res, 100, 200)), ("roughdielectric_ggx", material_ball(T.MATERIALS["roughdielectric_ggx"], res, res, 100, 200))),
                                                   (("all-fast", 256), ("hybrid", 128))):
    name = name + " " + mode
    g = api.Scene(ctx, d)
    rp = RenderParams(spp=spp, sampler="sobol", rfilter="box")
    fa, sa = g.render(rp, parity=True, flags=64); a = g.path_traces(spp)
    fb, sb = g.render(rp, parity=False, flags=64 | mflags); b = g.path_traces(spp)
    diff = a != b
    n = a.size
    x = (a ^ b)[diff]
    A, B = a[diff], b[diff]
    first = np.zeros(len(x), np.int64)
    for k in range(8):   # first differing byte
        pass
    byte = np.full(len(x), -1)
    for k in range(7, -1, -1):
        byte[((x >> np.uint64(8 * k)) & np.uint64(0xFF)) != 0] = k
    ea = ((A >> (8 * byte).astype(np.uint64)) & np.uint64(0xFF)).astype(np.int64)
    eb = ((B >> (8 * byte).astype(np.uint64)) & np.uint64(0xFF)).astype(np.int64)
    kinds = collections.Counter()
    prev_mat = collections.Counter()
    for i in range(len(x)):
        p, f = int(ea[i]), int(eb[i])
        if p == 0 or f == 0:
            k = "one build has no event here (earlier termination differs)"
        elif (p & 7) != (f & 7):
            k = f"hit material {p & 7} vs {f & 7}"
        elif ((p >> 4) & 3) != ((f >> 4) & 3):
            k = f"ending {(p >> 4) & 3} vs {(f >> 4) & 3} (0 continue, 1 roulette, 2 zero sample, 3 depth/miss) on material {p & 7}"
        elif (p & 0x40) != (f & 0x40):
            k = f"lobe: transmitted {bool(p & 0x40)} vs {bool(f & 0x40)} on material {p & 7}"
        elif (p & 8) != (f & 8):
            k = f"shadow ray emitted {bool(p & 8)} vs {bool(f & 8)} on material {p & 7}"
        else:
            k = "other"
        kinds[k] += 1
        if byte[i] > 0:   # what the previous vertex was (both builds agree on it)
            pe = int((A[i] >> np.uint64(8 * (byte[i] - 1))) & np.uint64(0xFF))
            prev_mat[f"bounce {int(byte[i]) + 1}, previous vertex on material {pe & 7}{' transmitted' if pe & 0x40 else ' reflected'}"] += 1
        else:
            prev_mat["bounce 1 (camera ray)"] += 1
    da, db = api.develop(fa).astype(np.float64), api.develop(fb).astype(np.float64)
    out[name] = {"rel_l2_vs_parity_build": float(np.sqrt(((da - db) ** 2).sum() / (da ** 2).sum())), "ms_parity_build": sa["ms_total"], "ms": sb["ms_total"],
                 "samples": int(n), "differing": int(diff.sum()), "fraction": float(diff.mean()), "first_difference": dict(kinds.most_common(20)),
                 "where": dict(prev_mat.most_common(20))}
    g.close()
print(json.dumps(out, indent=1))
