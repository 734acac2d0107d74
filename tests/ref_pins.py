"""Seeded inputs + one driver for the component functions that exist twice with the same C layout:
  coreref_*  the REFERENCE's own code (oracle/core_ref_shim.cpp -> oracle/_ref/libcoreref.so, compiled from /root/reference)
  orc_*      the oracle's restatement (oracle/mts_oracle.cpp)
tests/gen_golden.py runs the first and commits tests/golden/core_ref.npz; tests/test_oracle_reference_pins.py runs the second."""
import ctypes as C

import numpy as np

N = 400
MICROFACET = [(t, au, av, sv) for t in (0, 1, 2) for (au, av) in ((0.3, 0.3), (0.1, 0.4), (0.02, 0.02)) for sv in (1, 0) if not (t == 2 and sv)]
ETAS = (1.5, 1.0 / 1.5, 1.33, 1.0)
CONDUCTORS = (((0.2004, 0.9240, 1.1022), (3.9129, 2.4528, 2.1421)), ((0.1431, 0.3749, 1.4424), (3.9831, 2.3857, 1.6032)))
PMF_WEIGHTS = ([1.0, 2.0, 3.0, 4.0], [0.0, 1.0, 0.0, 0.0, 5.0, 0.0], [0.25] * 7 + [0.0], [3.0])


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dirs(rng, n, upper=False):
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if upper:
        d[:, 2] = np.abs(d[:, 2])
    return np.ascontiguousarray(d, np.float32)


def inputs(seed=31337):
    rng = np.random.default_rng(seed)
    x = {}
    x["wi_up"] = _dirs(rng, N, True)
    x["wi_up"][:8, 2] = np.float32(1e-3)  # grazing
    x["wi_up"][:8] /= np.linalg.norm(x["wi_up"][:8], axis=1, keepdims=True)
    x["m_up"] = _dirs(rng, N, True)
    x["samples"] = rng.random((N, 2)).astype(np.float32)
    x["samples"][:4] = np.array([[0, 0], [0.5, 0.5], [0.999999, 0.999999], [0, 0.999999]], np.float32)
    tris = rng.uniform(-1, 1, (N, 9)).astype(np.float32)
    tris[:6, 3:6] = tris[:6, 0:3]  # degenerate: two equal vertices
    tris[6:12, 2] = tris[6:12, 5] = tris[6:12, 8] = np.float32(0.25)  # axis-aligned
    x["tris"] = tris
    o = rng.uniform(-2, 2, (N, 3)).astype(np.float32)
    tgt = (tris[:, 0:3] * 0.3 + tris[:, 3:6] * 0.3 + tris[:, 6:9] * 0.4 + rng.normal(size=(N, 3)) * 0.3).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[12:20, 0] = 0.0  # rays parallel to an axis plane
    rays = np.concatenate([o, np.full((N, 1), 1e-4), d, np.full((N, 1), np.inf)], 1).astype(np.float32)
    rays[20:30, 7] = np.float32(0.5)  # short intervals
    x["rays"] = np.ascontiguousarray(rays)
    lo = rng.uniform(-1, 0, (N, 3)).astype(np.float32)
    x["boxes"] = np.ascontiguousarray(np.concatenate([lo, lo + rng.uniform(0.01, 1.5, (N, 3)).astype(np.float32)], 1))
    x["cos"] = np.concatenate([np.linspace(-1, 1, N - 6), [0.0, 1e-6, -1e-6, 1.0, -1.0, 0.5]]).astype(np.float32)
    x["dirs"] = _dirs(rng, N)
    x["normals"] = _dirs(rng, N)
    x["dpdu"] = (rng.normal(size=(N, 3)) * 2).astype(np.float32)
    x["pmf_samples"] = np.concatenate([rng.random(N - 4), [0.0, 0.999999, 0.5, 0.25]]).astype(np.float32)
    x["tea"] = rng.integers(0, 2 ** 32, (64, 2), dtype=np.uint64).astype(np.uint32)
    return x


def run(lib, prefix, x):
    """All component outputs as a dict of arrays; `lib` is a ctypes CDLL exporting <prefix><name> with the layouts of core_ref_shim.cpp."""
    fn = lambda name: getattr(lib, prefix + name)
    out = {}
    for (t, au, av, sv) in MICROFACET:
        key = f"mf_{t}_{au}_{av}_{sv}"
        a = np.zeros((N, 6), np.float32)
        b = np.zeros((N, 3), np.float32)
        # the oracle's historical entry points take uint64 counts; the new ones int -- both are fine with c_int for small n
        fn("microfacet_sample")(t, C.c_float(au), C.c_float(av), sv, C.c_uint64(N) if prefix == "orc_" else N, _f(x["wi_up"]), _f(x["samples"]), _f(a))
        fn("microfacet_eval")(t, C.c_float(au), C.c_float(av), sv, C.c_uint64(N) if prefix == "orc_" else N, _f(x["wi_up"]), _f(x["m_up"]), _f(b))
        out[key + "_sample"], out[key + "_eval"] = a, b
    rec = np.zeros((N, 12), np.uint32)
    st = np.zeros(N, np.int32)
    fn("triaccel_load")(N, _f(x["tris"]), rec.ctypes.data_as(C.POINTER(C.c_uint32)), st.ctypes.data_as(C.POINTER(C.c_int)))
    out["triaccel_records"], out["triaccel_status"] = rec, st
    a = np.zeros((N, 4), np.float32)
    fn("triaccel_intersect")(N, _f(x["tris"]), _f(x["rays"]), _f(a))
    out["triaccel_hits"] = a
    a = np.zeros((N, 3), np.float32)
    fn("aabb_intersect")(N, _f(x["boxes"]), _f(x["rays"]), _f(a))
    out["aabb"] = a
    for what in range(4):
        a = np.zeros((N, 3), np.float32)
        fn("warp")(what, N, _f(x["samples"]), _f(a))
        out[f"warp{what}"] = a
    for eta in ETAS:
        a = np.zeros((N, 2), np.float32)
        fn("fresnel_dielectric_ext")(N, _f(x["cos"]), C.c_float(eta), _f(a))
        out[f"fresnel_dielectric_{eta:.4f}"] = a
    for i, (eta, k) in enumerate(CONDUCTORS):
        a = np.zeros((N, 3), np.float32)
        e, kk = np.array(eta, np.float32), np.array(k, np.float32)
        fn("fresnel_conductor_exact_rgb")(N, _f(np.abs(x["cos"])), _f(e), _f(kk), _f(a))
        out[f"fresnel_conductor_{i}"] = a
    a = np.zeros((N, 3), np.float32)
    fn("reflect")(N, _f(x["dirs"]), _f(x["normals"]), _f(a))
    out["reflect"] = a
    for eta in (1.5, 1.33):
        a = np.zeros((N, 3), np.float32)
        ct = np.where(np.arange(N) % 2 == 0, -0.7, 0.6).astype(np.float32)
        fn("refract")(N, _f(x["dirs"]), _f(x["normals"]), C.c_float(eta), _f(ct), _f(a))
        out[f"refract_{eta}"] = a
    a = np.zeros((N, 6), np.float32)
    fn("coordinate_system")(N, _f(x["normals"]), _f(a))
    out["coordinate_system"] = a
    a = np.zeros((N, 9), np.float32)
    fn("shading_frame")(N, _f(x["normals"]), _f(x["dpdu"]), _f(a))
    out["shading_frame"] = a
    a = np.zeros((N, 3), np.float32)
    fn("triangle_sample")(N, _f(x["tris"]), _f(x["samples"]), _f(a))
    out["triangle_sample"] = a
    for i, w in enumerate(PMF_WEIGHTS):
        w = np.array(w, np.float32)
        idx, idx2 = np.zeros(N, np.uint32), np.zeros(N, np.uint32)
        reused, cdf = np.zeros(N, np.float32), np.zeros(len(w), np.float32)
        f = fn("pmf")
        f.restype = C.c_float
        total = f(len(w), _f(w), N, _f(x["pmf_samples"]), idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx2.ctypes.data_as(C.POINTER(C.c_uint32)), _f(reused), _f(cdf))
        out[f"pmf{i}_index"], out[f"pmf{i}_index_reuse"], out[f"pmf{i}_reused"], out[f"pmf{i}_pdf"], out[f"pmf{i}_total"] = idx, idx2, reused, cdf, np.float32(total)
    f = fn("tea")
    f.restype = C.c_uint64
    out["tea"] = np.array([[f(C.c_uint32(int(a)), C.c_uint32(int(b)), r) for r in (4, 8)] for a, b in x["tea"]], np.uint64)
    return out
