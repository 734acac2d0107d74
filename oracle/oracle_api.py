"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/_build/libmtsoracle.so (the CPU restatement of the Mitsuba 0.6 `path`
hot path).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under mitsuba_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_DATA = os.path.join(_HERE, "..", "mitsuba_b200", "data")


class OrcBsdf(C.Structure):
    _fields_ = [("type", C.c_int32), ("distr", C.c_int32), ("sampleVisible", C.c_int32), ("nested", C.c_int32),
                ("alphaU", C.c_float), ("alphaV", C.c_float), ("eta", C.c_float), ("thickness", C.c_float),
                ("reflectance", C.c_float * 3), ("transmittance", C.c_float * 3), ("etaC", C.c_float * 3),
                ("kC", C.c_float * 3), ("sigmaA", C.c_float * 3), ("nested2", C.c_int32), ("diffuseReflectance", C.c_float * 3),
                ("fdrInt", C.c_float), ("fdrExt", C.c_float), ("specSamplingWeight", C.c_float), ("nonlinear", C.c_int32),
                ("texture", C.c_int32)]


class OrcTextureDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32), ("filterType", C.c_int32),
                ("wrapU", C.c_int32), ("wrapV", C.c_int32), ("maxAnisotropy", C.c_float), ("uoffset", C.c_float),
                ("voffset", C.c_float), ("uscale", C.c_float), ("vscale", C.c_float)]


class OrcRenderParams(C.Structure):
    _fields_ = [("spp", C.c_int32), ("sampler", C.c_int32), ("seed", C.c_uint64), ("maxDepth", C.c_int32),
                ("rrDepth", C.c_int32), ("strictNormals", C.c_int32), ("hideEmitters", C.c_int32),
                ("rfilter", C.c_int32), ("rfilterParam", C.c_float), ("sampleLo", C.c_int32), ("sampleHi", C.c_int32),
                ("threads", C.c_int32), ("blockSize", C.c_int32), ("integrator", C.c_int32)]


class OrcMedium(C.Structure):
    _fields_ = [("type", C.c_int32), ("phase", C.c_int32), ("g", C.c_float), ("sigmaA", C.c_float * 3), ("sigmaS", C.c_float * 3),
                ("strategy", C.c_int32), ("samplingDensity", C.c_float), ("mediumSamplingWeight", C.c_float), ("scale", C.c_float),
                ("albedo", C.c_float * 3), ("res", C.c_int32 * 3), ("worldToGrid", C.c_float * 12), ("aabbMin", C.c_float * 3),
                ("aabbMax", C.c_float * 3), ("density", C.POINTER(C.c_float))]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "rays", "shadowRays", "pathLengthSum", "nodeVisits",
                                          "primTests", "badSamples", "dimOverflow")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force=False):
    so = os.path.join(_HERE, "_build", "libmtsoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mts_oracle.cpp", "orc_math.h", "orc_sampler.h", "orc_accel.h", "orc_bsdf.h", "orc_medium.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_scene_new.restype = C.c_void_p
        L.orc_add_bsdf.restype = C.c_int
        L.orc_add_mesh.restype = C.c_int
        L.orc_add_medium.restype = C.c_int
        L.orc_add_texture.restype = C.c_int
        L.orc_tea.restype = C.c_uint64
        L.orc_bsdf_type.restype = C.c_uint32
        L.orc_hardware_threads.restype = C.c_int
    return _LIB


_TABLES = None


def sobol_tables():
    global _TABLES
    if _TABLES is None:
        m32 = np.fromfile(os.path.join(_DATA, "sobol_matrices32.bin"), dtype="<u4")
        vdc = np.fromfile(os.path.join(_DATA, "sobol_vdc.bin"), dtype="<u8")
        inv = np.fromfile(os.path.join(_DATA, "sobol_vdc_inv.bin"), dtype="<u8")
        _TABLES = (np.ascontiguousarray(m32), np.ascontiguousarray(vdc), np.ascontiguousarray(inv))
    return _TABLES


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def make_bsdf_array(flat_list):
    arr = (OrcBsdf * max(1, len(flat_list)))()
    for i, d in enumerate(flat_list):
        b = arr[i]
        b.type, b.distr, b.sampleVisible, b.nested = d["type"], d["distr"], d["sampleVisible"], d["nested"]
        b.alphaU, b.alphaV, b.eta, b.thickness = d["alphaU"], d["alphaV"], d["eta"], d["thickness"]
        for k in ("reflectance", "transmittance", "etaC", "kC", "sigmaA", "diffuseReflectance"):
            for j in range(3):
                getattr(b, k)[j] = d[k][j]
        b.nested2, b.fdrInt, b.fdrExt, b.specSamplingWeight, b.nonlinear = d["nested2"], d["fdrInt"], d["fdrExt"], d["specSamplingWeight"], d["nonlinear"]
        b.texture = d.get("texture", -1)
    return arr


def make_texture_desc(d):
    t = OrcTextureDesc()
    for k, _ in OrcTextureDesc._fields_:
        setattr(t, k, d[k])
    return t


def instance_matrices(to_world):
    """float32 (M, M^-1) of an instance transform; the inverse is taken in float64 once so that both sides use the same numbers."""
    M64 = np.asarray(to_world, np.float64)
    return np.ascontiguousarray(M64, np.float32), np.ascontiguousarray(np.linalg.inv(M64), np.float32)


def make_medium(d):
    m = OrcMedium()
    m.type, m.phase, m.g, m.strategy = d["type"], d["phase"], d["g"], d["strategy"]
    m.samplingDensity, m.mediumSamplingWeight, m.scale = d["samplingDensity"], d["mediumSamplingWeight"], d["scale"]
    for k in ("sigmaA", "sigmaS", "albedo", "aabbMin", "aabbMax"):
        for j in range(3):
            getattr(m, k)[j] = d[k][j]
    for j in range(3):
        m.res[j] = d["res"][j]
    for j in range(12):
        m.worldToGrid[j] = d["worldToGrid"][j]
    m.density = d["density"].ctypes.data_as(C.POINTER(C.c_float)) if d["density"] is not None else None
    return m


SAMPLERS = {"sobol": 0, "independent_sfmt": 1, "independent": 2}
RFILTERS = {"box": 0, "gaussian": 1}


def make_params(rp, threads=0, sampler=None):
    p = OrcRenderParams()
    p.spp = rp.spp
    p.sampler = SAMPLERS[sampler or rp.sampler]
    p.seed = rp.seed
    p.maxDepth, p.rrDepth = rp.max_depth, rp.rr_depth
    p.strictNormals, p.hideEmitters = int(rp.strict_normals), int(rp.hide_emitters)
    p.rfilter, p.rfilterParam = RFILTERS[rp.rfilter], rp.rfilter_param
    p.sampleLo, p.sampleHi = rp.sample_lo, rp.sample_hi
    p.threads, p.blockSize = threads, 32
    p.integrator = {"path": 0, "volpath": 1}[getattr(rp, "integrator", "path")]
    return p


class OracleScene:
    """Oracle-side scene built from a mitsuba_b200.scene.SceneDesc (plain data only)."""

    def __init__(self, desc, use_tree=True, sample_to_camera=None, instance_inverses=None):
        """sample_to_camera / instance_inverses: host-side matrix set-up handed in from elsewhere (the image-level pins pass the
        reference's own Transform::inverse() results so that both renderers start from the same numbers)."""
        L = lib()
        self.L = L
        self.h = C.c_void_p(L.orc_scene_new())
        self._keep = []
        m32, vdc, inv = sobol_tables()
        L.orc_set_sobol_tables(self.h, _p(m32, C.c_uint32), _p(vdc, C.c_uint64), _p(inv, C.c_uint64))
        flat, ids = desc.flat_bsdfs()
        self.flat_bsdfs = flat
        arr = make_bsdf_array(flat)
        for i in range(len(flat)):
            L.orc_add_bsdf(self.h, C.byref(arr[i]))
        self.flat_textures = desc.flat_textures() if hasattr(desc, "flat_textures") else []
        for d in self.flat_textures:
            L.orc_add_texture(self.h, C.byref(make_texture_desc(d)), _p(d["pixels"]))
        for _ in range(desc.n_groups() if hasattr(desc, "n_groups") else 0):
            L.orc_add_shapegroup(self.h)
        for m, bid in zip(desc.meshes, ids):
            P = np.ascontiguousarray(m.P, np.float32)
            N = np.ascontiguousarray(m.N, np.float32) if m.N is not None else None
            UV = np.ascontiguousarray(m.UV, np.float32) if m.UV is not None else None
            I = np.ascontiguousarray(m.idx, np.uint32)
            rad = np.asarray(m.radiance, np.float32) if m.radiance is not None else None
            mid = L.orc_add_mesh(self.h, _p(P), _p(N), _p(UV), C.c_uint32(len(P)), _p(I, C.c_uint32), C.c_uint32(len(I)),
                                 C.c_int(bid), _p(rad), C.c_float(m.sampling_weight))
            if getattr(m, "group", -1) >= 0:
                if m.radiance is not None:
                    raise ValueError("Instancing of emitters is not supported")  # shapegroup.cpp:115-116
                L.orc_set_mesh_group(self.h, C.c_int(mid), C.c_int(m.group))
        for k, inst in enumerate(getattr(desc, "instances", [])):
            M, Minv = instance_matrices(inst.to_world)
            if instance_inverses is not None:
                Minv = np.ascontiguousarray(instance_inverses[k], np.float32)
            L.orc_add_instance(self.h, C.c_int(inst.group), _p(M), _p(Minv))
        if getattr(desc, "env_radiance", None) is not None:
            rad = np.asarray(desc.env_radiance, np.float32)
            L.orc_add_constant_emitter(self.h, _p(rad), C.c_float(desc.env_sampling_weight))
        if getattr(desc, "envmap", None) is not None:
            if getattr(desc, "env_radiance", None) is not None:
                raise ValueError("The scene may only contain one environment emitter")  # scene.cpp:510-514
            em = desc.envmap
            px = np.ascontiguousarray(em.pixels, np.float32)
            M, Minv = em.matrices()
            L.orc_add_envmap_emitter.restype = C.c_int
            if L.orc_add_envmap_emitter(self.h, C.c_int(px.shape[1]), C.c_int(px.shape[0]), _p(px), C.c_float(em.scale), _p(M), _p(Minv), C.c_float(em.sampling_weight)) < 0:
                raise ValueError("The environment map is completely black or holds a non-finite value")  # envmap.cpp:311-315
        media, mids = desc.flat_media()
        self.flat_media = [md.flat() for md in media]
        for d in self.flat_media:
            L.orc_add_medium(self.h, C.byref(make_medium(d)))
        for i, (mi, me) in enumerate(mids):
            if mi >= 0 or me >= 0:
                L.orc_set_mesh_media(self.h, C.c_int(i), C.c_int(mi), C.c_int(me))
        cam = desc.camera
        self.W, self.H = cam.film_size()   # the crop window is the film the integrator sees (Film::getCropSize)
        c2w = np.ascontiguousarray(cam.to_world, np.float32)
        # the camera matrix may be handed over from the implementation under test so both sides start from
        # identical float32 inputs (its derivation is host-side set-up, perspective.cpp:146-153, not the hot path)
        s2c = np.ascontiguousarray(cam.sample_to_camera() if sample_to_camera is None else sample_to_camera, np.float32)
        L.orc_set_camera(self.h, _p(c2w), _p(s2c), C.c_float(cam.near), C.c_float(cam.far), C.c_int(self.W), C.c_int(self.H))
        if getattr(cam, "aperture_radius", 0.0) > 0:
            L.orc_set_thinlens(self.h, C.c_float(cam.aperture_radius), C.c_float(cam.focus_distance if cam.focus_distance > 0 else cam.far))
        L.orc_commit(self.h, C.c_int(1 if use_tree else 0))

    def __del__(self):
        try:
            self.L.orc_scene_free(self.h)
        except Exception:
            pass

    def accel_info(self):
        out = np.zeros(5, np.uint64); bb = np.zeros(6, np.float32)
        self.L.orc_accel_info(self.h, _p(out, C.c_uint64), _p(bb))
        return dict(n_tri=int(out[0]), n_nodes=int(out[1]), n_indices=int(out[2]), n_leaves=int(out[3]),
                    max_depth=int(out[4]), aabb=bb)

    def triaccel(self):
        n = self.accel_info()["n_tri"]
        buf = np.zeros((n, 12), np.float32)
        self.L.orc_get_triaccel(self.h, buf.ctypes.data_as(C.c_void_p))
        return buf

    def trace(self, rays, mode=0, accel=-1):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = len(rays)
        t = np.zeros(n, np.float32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32); prim = np.zeros(n, np.uint32)
        self.L.orc_trace(self.h, C.c_uint64(n), _p(rays), C.c_int(mode), C.c_int(accel), _p(t), _p(u), _p(v), _p(prim, C.c_uint32))
        return t, u, v, prim

    def camera_rays(self, pos):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        rays = np.zeros((len(pos), 8), np.float32)
        self.L.orc_camera_rays(self.h, C.c_uint64(len(pos)), _p(pos), _p(rays))
        return rays

    def intersect_full(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros((len(rays), 24), np.float32)
        self.L.orc_intersect_full(self.h, C.c_uint64(len(rays)), _p(rays), _p(out))
        return out

    # ---- bitmap textures (orc_texture.h) ----
    def texture_info(self, tex):
        out = np.zeros(1 + 2 * 32, np.int32)
        mx, sc = C.c_float(), C.c_float()
        self.L.orc_texture_info(self.h, C.c_int(tex), _p(out, C.c_int32), C.byref(mx), C.byref(sc))
        n = int(out[0])
        return dict(levels=n, sizes=[(int(out[1 + 2 * l]), int(out[2 + 2 * l])) for l in range(n)], maximum=mx.value, bsdf_scale=sc.value)

    def texture_level(self, tex, level):
        info = self.texture_info(tex)
        w, h = info["sizes"][level]
        ch = self.flat_textures[tex]["channels"]
        out = np.zeros((h, w, ch), np.float32)
        self.L.orc_texture_level(self.h, C.c_int(tex), C.c_int(level), _p(out))
        return out

    def texture_eval(self, tex, uv, partials=None):
        """Texture2D::eval: uv (n,2); partials (n,4) = dudx, dudy, dvdx, dvdy for the filtered look-up or None."""
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        pt = np.ascontiguousarray(partials, np.float32).reshape(-1, 4) if partials is not None else None
        out = np.zeros((len(uv), 3), np.float32)
        self.L.orc_texture_eval(self.h, C.c_int(tex), C.c_uint64(len(uv)), _p(uv), _p(pt), _p(out))
        return out

    def primary_partials(self, pos, spp):
        """(n,8): valid, u, v, dudx, dudy, dvdx, dvdy, mesh for camera rays through film positions `pos`."""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
        out = np.zeros((len(pos), 8), np.float32)
        self.L.orc_primary_partials(self.h, C.c_uint64(len(pos)), _p(pos), C.c_int(spp), _p(out))
        return out

    def sample_emitter_direct(self, ref, samples):
        ref = np.ascontiguousarray(ref, np.float32).reshape(-1, 6)
        samples = np.ascontiguousarray(samples, np.float32).reshape(-1, 2)
        out = np.zeros((len(ref), 12), np.float32)
        self.L.orc_sample_emitter_direct(self.h, C.c_uint64(len(ref)), _p(ref), _p(samples), _p(out))
        return out

    def sampler_stream(self, kind, seed, spp, px, py, sample_idx, ndim):
        out = np.zeros(ndim, np.float32)
        self.L.orc_sampler_stream(self.h, C.c_int(SAMPLERS[kind]), C.c_uint64(seed), C.c_int(self.W), C.c_int(self.H),
                                  C.c_int(spp), C.c_int(px), C.c_int(py), C.c_int(sample_idx), C.c_int(ndim), _p(out))
        return out

    # ---- medium component probes (tests/test_oracle_volpath.py, tests/test_gpu_volpath.py) ----
    def medium_transmittance(self, medium, rays, seed=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros((len(rays), 3), np.float32)
        self.L.orc_medium_transmittance(self.h, C.c_int(medium), C.c_uint64(len(rays)), _p(rays), C.c_uint64(seed), _p(out))
        return out

    def medium_sample_distance(self, medium, rays, seed=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.zeros((len(rays), 12), np.float32)
        self.L.orc_medium_sample_distance(self.h, C.c_int(medium), C.c_uint64(len(rays)), _p(rays), C.c_uint64(seed), _p(out))
        return out

    def medium_density(self, medium, p):
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3)
        out = np.zeros(len(p), np.float32)
        self.L.orc_medium_density(self.h, C.c_int(medium), C.c_uint64(len(p)), _p(p), _p(out))
        return out

    def phase(self, medium, wi, samples):
        wi = np.ascontiguousarray(wi, np.float32).reshape(-1, 3)
        samples = np.ascontiguousarray(samples, np.float32).reshape(-1, 2)
        out = np.zeros((len(wi), 5), np.float32)
        self.L.orc_phase(self.h, C.c_int(medium), C.c_uint64(len(wi)), _p(wi), _p(samples), _p(out))
        return out

    def render(self, rp, threads=0, sampler=None, per_sample=False):
        p = make_params(rp, threads, sampler)
        film = np.zeros((self.H, self.W, 5), np.float32)
        st = OrcStats()
        if per_sample:
            hi = rp.sample_hi if rp.sample_hi > 0 else rp.spp
            ps = np.zeros((self.H, self.W, hi - rp.sample_lo, 4), np.float32)
            self.L.orc_render_samples(self.h, C.byref(p), _p(film), C.byref(st), _p(ps))
            return film, st.as_dict(), ps
        self.L.orc_render(self.h, C.byref(p), _p(film), C.byref(st))
        return film, st.as_dict()


def develop(film):
    film = np.ascontiguousarray(film, np.float32)
    H, W = film.shape[:2]
    rgb = np.zeros((H, W, 3), np.float32)
    lib().orc_develop(_p(film), C.c_int(W), C.c_int(H), _p(rgb))
    return rgb


def filter_table(kind, param):
    v = np.zeros(32, np.float32); r = C.c_float(); b = C.c_int()
    lib().orc_filter_table(C.c_int(RFILTERS[kind]), C.c_float(param), _p(v), C.byref(r), C.byref(b))
    return v, r.value, b.value


def splat(W, H, kind, param, pos, val):
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 2)
    val = np.ascontiguousarray(val, np.float32).reshape(-1, 4)
    film = np.zeros((H, W, 5), np.float32)
    lib().orc_splat(C.c_int(W), C.c_int(H), C.c_int(RFILTERS[kind]), C.c_float(param), C.c_uint64(len(pos)), _p(pos), _p(val), _p(film))
    return film


def sfmt_words(seed, n):
    out = np.zeros(n, np.uint64)
    lib().orc_sfmt_words(C.c_uint64(seed), C.c_uint64(n), _p(out, C.c_uint64))
    return out


def sfmt_floats(seed, n):
    out = np.zeros(n, np.float32)
    lib().orc_sfmt_floats(C.c_uint64(seed), C.c_uint64(n), _p(out))
    return out


def tea(v0, v1, rounds=4):
    return int(lib().orc_tea(C.c_uint32(v0), C.c_uint32(v1), C.c_int(rounds)))


def sobol_sample(index, dim, scramble=0):
    m32, _, _ = sobol_tables()
    index = np.ascontiguousarray(index, np.uint64); dim = np.ascontiguousarray(np.broadcast_to(dim, index.shape), np.uint32)
    out = np.zeros(index.shape, np.float32)
    lib().orc_sobol_sample(_p(m32, C.c_uint32), C.c_uint64(index.size), _p(index, C.c_uint64), _p(dim, C.c_uint32), C.c_uint32(scramble), _p(out))
    return out


def sobol_lookup(m, frame, px, py, scramble=0):
    _, vdc, inv = sobol_tables()
    frame = np.ascontiguousarray(frame, np.uint32)
    px = np.ascontiguousarray(np.broadcast_to(px, frame.shape), np.uint32)
    py = np.ascontiguousarray(np.broadcast_to(py, frame.shape), np.uint32)
    out = np.zeros(frame.shape, np.uint64)
    lib().orc_sobol_lookup(_p(vdc, C.c_uint64), _p(inv, C.c_uint64), C.c_uint32(m), C.c_uint64(frame.size), _p(frame, C.c_uint32),
                           _p(px, C.c_uint32), _p(py, C.c_uint32), C.c_uint64(scramble), _p(out, C.c_uint64))
    return out


def bsdf_eval(flat_list, bid, wi, wo, discrete=False):
    arr = make_bsdf_array(flat_list)
    wi = np.ascontiguousarray(wi, np.float32).reshape(-1, 3); wo = np.ascontiguousarray(wo, np.float32).reshape(-1, 3)
    rgb = np.zeros((len(wi), 3), np.float32); pdf = np.zeros(len(wi), np.float32)
    (lib().orc_bsdf_eval_discrete if discrete else lib().orc_bsdf_eval)(arr, C.c_int(len(flat_list)), C.c_int(bid), C.c_uint64(len(wi)), _p(wi), _p(wo), _p(rgb), _p(pdf))
    return rgb, pdf


def bsdf_sample(flat_list, bid, wi, samples):
    """samples: n x 3 (2D sample, extra 1D).  Returns dict of wo, weight, pdf, type, eta."""
    arr = make_bsdf_array(flat_list)
    wi = np.ascontiguousarray(wi, np.float32).reshape(-1, 3); samples = np.ascontiguousarray(samples, np.float32).reshape(-1, 3)
    out = np.zeros((len(wi), 10), np.float32)
    lib().orc_bsdf_sample(arr, C.c_int(len(flat_list)), C.c_int(bid), C.c_uint64(len(wi)), _p(wi), _p(samples), _p(out))
    return dict(wo=out[:, 0:3], weight=out[:, 3:6], pdf=out[:, 6], type=out[:, 7].astype(np.uint32), eta=out[:, 8])


def bsdf_type(flat_list, bid):
    arr = make_bsdf_array(flat_list)
    return int(lib().orc_bsdf_type(arr, C.c_int(len(flat_list)), C.c_int(bid)))


def microfacet_sample(distr, alpha_u, alpha_v, sample_visible, wi, samples):
    wi = np.ascontiguousarray(wi, np.float32).reshape(-1, 3); samples = np.ascontiguousarray(samples, np.float32).reshape(-1, 2)
    out = np.zeros((len(wi), 6), np.float32)
    lib().orc_microfacet_sample(C.c_int(distr), C.c_float(alpha_u), C.c_float(alpha_v), C.c_int(int(sample_visible)),
                                C.c_uint64(len(wi)), _p(wi), _p(samples), _p(out))
    return out


def microfacet_eval(distr, alpha_u, alpha_v, sample_visible, wi, m):
    wi = np.ascontiguousarray(wi, np.float32).reshape(-1, 3); m = np.ascontiguousarray(m, np.float32).reshape(-1, 3)
    out = np.zeros((len(wi), 3), np.float32)
    lib().orc_microfacet_eval(C.c_int(distr), C.c_float(alpha_u), C.c_float(alpha_v), C.c_int(int(sample_visible)),
                              C.c_uint64(len(wi)), _p(wi), _p(m), _p(out))
    return out


def hardware_threads():
    return int(lib().orc_hardware_threads())
