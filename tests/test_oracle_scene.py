"""Scene-level self-checks of the oracle (the pieces SURVEY.md 8c lists as "not pinned by the reference")."""
import dataclasses

import numpy as np
import pytest

from mitsuba_b200.scene import Bsdf, Camera, Mesh, RenderParams, SceneDesc, cornell_box, look_at, material_ball, uv_sphere
from oracle import oracle_api as O


@pytest.fixture(scope="module")
def cbox():
    d = cornell_box(48, 48)
    return d, O.OracleScene(d)


def random_rays(rng, n, lo, hi):
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, np.full((n, 1), 1e-4, np.float32), d, np.full((n, 1), np.inf, np.float32)], 1).astype(np.float32)


def test_cornell_geometry(cbox):
    d, sc = cbox
    assert d.n_triangles() == 32
    info = sc.accel_info()
    assert info["n_tri"] == 32 and info["max_depth"] == int(8 + 1.3 * 5)
    # every wall/box normal faces the room interior / outward (one-sided diffuse BSDFs)
    c = np.array([278, 274, 280], np.float32)
    for m in d.meshes:
        P, I = m.P, m.idx
        n = np.cross(P[I[:, 1]] - P[I[:, 0]], P[I[:, 2]] - P[I[:, 0]])
        cen = P[I].mean(1)
        s = (n * (c - cen)).sum(1)
        if m.name in ("short", "tall"):
            inner = P.mean(0)
            assert ((n * (cen - inner)).sum(1) > 0).all()
        else:
            assert (s > 0).all(), m.name


def test_kdtree_equals_brute_force(cbox):
    """Havran traversal over the SAH tree returns the same argmin-t TriAccel hit as testing every triangle."""
    _, sc = cbox
    rays = random_rays(np.random.default_rng(3), 30000, 5, 550)
    t0, u0, v0, p0 = sc.trace(rays, 0, accel=0)
    t1, u1, v1, p1 = sc.trace(rays, 0, accel=1)
    assert np.array_equal(p0, p1) and np.array_equal(t0, t1) and np.array_equal(u0, u1) and np.array_equal(v0, v1)
    rays[:, 7] = np.random.default_rng(4).uniform(20, 700, len(rays))
    assert np.array_equal(sc.trace(rays, 1, accel=0)[3], sc.trace(rays, 1, accel=1)[3])


def test_kdtree_equals_brute_force_mesh():
    P, N, _, I = uv_sphere((0, 0, 0), 1.0, 24, 48)
    d = SceneDesc([Mesh(P, I, N=N, bsdf=Bsdf("diffuse"))], Camera(look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), width=16, height=16))
    sc = O.OracleScene(d)
    rays = random_rays(np.random.default_rng(5), 20000, -2, 2)
    a = sc.trace(rays, 0, accel=0); b = sc.trace(rays, 0, accel=1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[3] != 0xFFFFFFFF).mean() > 0.1


def test_triaccel_matches_moller_trumbore(cbox):
    """triaccel.h:96-158 hit distances / barycentrics agree with a float64 Moller-Trumbore on the hit triangle."""
    d, sc = cbox
    rays = random_rays(np.random.default_rng(6), 5000, 20, 530)
    t, u, v, prim = sc.trace(rays, 0)
    tris = np.concatenate([m.P[m.idx] for m in d.meshes]).astype(np.float64)
    ok = prim != 0xFFFFFFFF
    A = tris[prim[ok]]
    o = rays[ok, :3].astype(np.float64); dd = rays[ok, 4:7].astype(np.float64)
    e1, e2 = A[:, 1] - A[:, 0], A[:, 2] - A[:, 0]
    pv = np.cross(dd, e2); det = (e1 * pv).sum(1); tv = o - A[:, 0]
    uu = (tv * pv).sum(1) / det; qv = np.cross(tv, e1); vv = (dd * qv).sum(1) / det; tt = (e2 * qv).sum(1) / det
    assert np.allclose(t[ok], tt, rtol=2e-4) and np.allclose(u[ok], uu, atol=2e-4) and np.allclose(v[ok], vv, atol=2e-4)


def test_intersection_record(cbox):
    """skdtree.h:343-428: barycentric p lies on the ray, frames are orthonormal, wi = toLocal(-d)."""
    _, sc = cbox
    rays = random_rays(np.random.default_rng(7), 2000, 20, 530)
    rec = sc.intersect_full(rays)
    ok = rec[:, 21] == 1
    p, gn, sn, s, t, wi, tt = rec[ok, 0:3], rec[ok, 3:6], rec[ok, 6:9], rec[ok, 9:12], rec[ok, 12:15], rec[ok, 15:18], rec[ok, 18]
    assert np.allclose(p, rays[ok, :3] + tt[:, None] * rays[ok, 4:7], atol=2e-2)
    for a, b in ((sn, s), (sn, t), (s, t)):
        assert np.abs((a * b).sum(1)).max() < 1e-5
    assert np.allclose(np.linalg.norm(sn, axis=1), 1, atol=1e-5)
    assert np.allclose(wi[:, 2], -(rays[ok, 4:7] * sn).sum(1), atol=1e-5)
    assert np.allclose(gn, sn)   # no vertex normals: shading normal = face normal


def test_emitter_direct_sampling(cbox):
    """area.cpp:158-183 / shape.cpp:102-126: value * pdf = radiance, solid-angle pdf = dist^2 / (A |cos|)."""
    _, sc = cbox
    rng = np.random.default_rng(8)
    ref = np.concatenate([rng.uniform(50, 500, (4000, 3)) * [1, 0.6, 1], np.zeros((4000, 3))], 1).astype(np.float32)
    out = sc.sample_emitter_direct(ref, rng.uniform(size=(4000, 2)))
    vis = out[:, 8] == 1
    assert vis.mean() > 0.5
    d, dist, pdf, val, p = out[vis, 0:3], out[vis, 3], out[vis, 4], out[vis, 5:8], out[vis, 9:12]
    assert np.allclose(val * pdf[:, None], [17, 12, 4], rtol=1e-4)
    assert np.allclose(p[:, 1], 548.3) and (p[:, 0] >= 213).all() and (p[:, 0] <= 343).all() and (p[:, 2] >= 227).all() and (p[:, 2] <= 332).all()
    area = 130 * 105
    assert np.allclose(pdf, dist ** 2 / (area * np.abs(d[:, 1])), rtol=1e-4)


def test_film_weight_channel(cbox):
    """The weight channel equals the sum of filter weights: spp * norm^2 for interior pixels of a gaussian film."""
    _, sc = cbox
    film, st = sc.render(RenderParams(spp=8, rfilter="gaussian"), threads=2)
    assert st["samples"] == 48 * 48 * 8 and st["badSamples"] == 0
    w = film[8:-8, 8:-8, 4]
    assert abs(w.mean() / 8 - 1) < 0.05
    assert np.allclose(film[12:36, 12:36, 3], film[12:36, 12:36, 4], rtol=1e-5)   # alpha == weight where every camera ray hits


def test_render_is_deterministic_and_thread_independent(cbox):
    _, sc = cbox
    rp = RenderParams(spp=4, rfilter="box")
    f1, s1 = sc.render(rp, threads=1)
    f2, s2 = sc.render(rp, threads=4)
    assert np.array_equal(f1, f2) and s1 == s2


def test_sample_ranges_add_up(cbox):
    """Sharding mirror: films of [0,3) and [3,8) add up to the film of [0,8) (Sobol' state is a pure function)."""
    _, sc = cbox
    rp = RenderParams(spp=8, rfilter="gaussian")
    full, _ = sc.render(rp)
    a, _ = sc.render(dataclasses.replace(rp, sample_lo=0, sample_hi=3))
    b, _ = sc.render(dataclasses.replace(rp, sample_lo=3, sample_hi=8))
    assert np.allclose(a + b, full, rtol=1e-5, atol=1e-6)
    assert np.array_equal((a + b)[..., 4] > 0, full[..., 4] > 0)


def test_radiance_linearity(cbox):
    """Li is linear in the emitted radiance: scaling it by 2 scales the film's RGB exactly (power of two)."""
    d, sc = cbox
    d2 = cornell_box(48, 48)
    d2.meshes[-1].radiance = tuple(2 * x for x in d2.meshes[-1].radiance)
    rp = RenderParams(spp=4, rfilter="box")
    f1, _ = sc.render(rp); f2, _ = O.OracleScene(d2).render(rp)
    assert np.array_equal(f2[..., :3], 2 * f1[..., :3]) and np.array_equal(f2[..., 3:], f1[..., 3:])


def test_white_furnace():
    """A diffuse sphere (albedo rho) seen from inside a uniformly emitting... closed box emitter: path.cpp's NEE+BSDF
    MIS estimate of an enclosed diffuse patch converges to L * rho / (1 - ... ) -- here the simplest closed form:
    camera inside an emitting sphere of radiance L with a black (absorbing) BSDF sees exactly L."""
    P, N, _, I = uv_sphere((0, 0, 0), 5.0, 16, 32, smooth=False)
    I = I[:, ::-1].copy()   # normals inward
    d = SceneDesc([Mesh(P, I, bsdf=Bsdf("diffuse", reflectance=(0, 0, 0)), radiance=(3.0, 2.0, 1.0))],
                  Camera(look_at((0, 0, 0), (0, 0, 1), (0, 1, 0)), fov=60, near=0.01, far=100, width=24, height=24))
    film, st = O.OracleScene(d).render(RenderParams(spp=4, rfilter="box"))
    rgb = O.develop(film)
    assert np.allclose(rgb, [3, 2, 1], rtol=1e-5)
    assert st["pathLengthSum"] == st["samples"]    # absorbing BSDF: every path has length 1


def test_max_depth_semantics(cbox):
    """integrator.cpp:195-199: maxDepth=1 shows only directly visible emitters, 2 adds single-bounce direct light."""
    _, sc = cbox
    f1, s1 = sc.render(RenderParams(spp=2, rfilter="box", max_depth=1))
    rgb1 = O.develop(f1)
    lit = rgb1.sum(2) > 0
    assert 0 < lit.mean() < 0.05 and np.allclose(rgb1.max((0, 1)), [17, 12, 4]) and (rgb1 <= np.float32([17, 12, 4]) * 1.0001).all()
    assert s1["pathLengthSum"] == s1["samples"] and s1["shadowRays"] == 0
    f2, s2 = sc.render(RenderParams(spp=2, rfilter="box", max_depth=2))
    assert (O.develop(f2).sum(2) > 0).mean() > 0.5 and s2["pathLengthSum"] <= 2 * s2["samples"]


@pytest.mark.parametrize("bsdf", [Bsdf("roughconductor", distribution="ggx", alpha_u=0.1, alpha_v=0.1, eta=(0.2, 0.92, 1.1), k=(3.9, 2.45, 2.14)),
                                  Bsdf("roughdielectric", distribution="ggx", alpha_u=0.1, alpha_v=0.1, int_ior="bk7", ext_ior="air"),
                                  Bsdf("coating", int_ior=1.5, ext_ior=1.0, nested=Bsdf("diffuse", reflectance=(0.6, 0.2, 0.2)))])
def test_material_ball_renders(bsdf):
    d = material_ball(bsdf, 32, 32, 24, 48)
    film, st = O.OracleScene(d).render(RenderParams(spp=8, rfilter="gaussian"))
    rgb = O.develop(film)
    assert np.isfinite(rgb).all() and st["badSamples"] == 0 and rgb.mean() > 0.01 and st["dimOverflow"] == 0


def test_splat_matches_block_put():
    """ImageBlock::put through 32x32 blocks + film merge (imageblock.h:103-204) against an independent numpy
    transcription.  The footprint arithmetic is block relative (pos - 0.5 - (offset - border)): the filter bin a
    pixel falls into can differ by one from full-frame arithmetic because the float32 rounding differs."""
    rng = np.random.default_rng(9)
    W, H, n = 70, 45, 4000
    pos = rng.uniform(0, [W, H], (n, 2)).astype(np.float32)
    pos[:50] = np.floor(pos[:50])                        # exact pixel corners: the box filter's 2x2 footprint case
    val = rng.uniform(0, 2, (n, 4)).astype(np.float32)
    f32 = np.float32
    for kind, param in (("box", 0.5), ("gaussian", 0.5), ("gaussian", 0.8)):
        film = O.splat(W, H, kind, param, pos, val)
        tab, radius, border = O.filter_table(kind, param)
        radius = f32(radius); sf = f32(31) / radius
        ref = np.zeros((H, W, 5), np.float64)
        for (x, y), v in zip(pos, val):
            ox, oy = (int(np.floor(x)) // 32) * 32, (int(np.floor(y)) // 32) * 32
            sx, sy = min(32, W - ox) + 2 * border, min(32, H - oy) + 2 * border
            px, py = f32(f32(x - f32(0.5)) - f32(ox - border)), f32(f32(y - f32(0.5)) - f32(oy - border))
            xs = range(max(int(np.ceil(px - radius)), 0), min(int(np.floor(px + radius)), sx - 1) + 1)
            ys = range(max(int(np.ceil(py - radius)), 0), min(int(np.floor(py + radius)), sy - 1) + 1)
            for yy in ys:
                wy = tab[min(int(abs(f32(f32(yy) - py) * sf)), 31)]
                for xx in xs:
                    wx = tab[min(int(abs(f32(f32(xx) - px) * sf)), 31)]
                    fx, fy = ox - border + xx, oy - border + yy
                    if 0 <= fx < W and 0 <= fy < H:
                        ref[fy, fx] += f32(wx * wy) * np.array([v[0], v[1], v[2], v[3], 1.0])
        assert np.allclose(film, ref, rtol=2e-4, atol=1e-5), kind   # f32 accumulation order vs f64


def test_constant_environment_white_furnace():
    """`constant` emitter (constant.cpp): a white diffuse sphere under a uniform sky shows exactly the sky radiance."""
    P, N, _, I = uv_sphere((0, 0, 0), 1.0, 32, 64)
    d = SceneDesc([Mesh(P, I, N=N, bsdf=Bsdf("diffuse", reflectance=(1, 1, 1)))], Camera(look_at((0, 0, -4), (0, 0, 0), (0, 1, 0)), fov=40, width=24, height=24),
                  env_radiance=(0.7, 0.8, 0.9))
    sc = O.OracleScene(d)
    film, st = sc.render(RenderParams(spp=128, sampler="independent", rfilter="box", rr_depth=50))
    rgb = O.develop(film)
    assert np.allclose(rgb.reshape(-1, 3).mean(0), (0.7, 0.8, 0.9), rtol=2e-3)
    assert rgb.reshape(-1, 3).std(0).max() < 5e-3
    # hideEmitters hides the directly visible sky only
    film, _ = sc.render(RenderParams(spp=16, sampler="independent", rfilter="box", hide_emitters=True))
    rgb = O.develop(film)
    assert rgb[0, 0].max() == 0 and rgb[12, 12].min() > 0.3
    # grey sphere: radiance = L * rho / (1 - 0) first bounce only sees sky -> rho * L at the centre... the full series: L * rho (no interreflection on a convex body)
    d.meshes[0].bsdf = Bsdf("diffuse", reflectance=(0.5, 0.5, 0.5))
    film, _ = O.OracleScene(d).render(RenderParams(spp=256, sampler="independent", rfilter="box"))
    assert np.allclose(O.develop(film)[12, 12], np.float32([0.7, 0.8, 0.9]) * 0.5, rtol=0.02)


def test_instancing_equals_flattened_geometry():
    """shapegroup + instance (src/shapes/{shapegroup,instance}.cpp) against the same geometry transformed into world space."""
    from mitsuba_b200.scene import stress_scene
    rp = RenderParams(spp=8, sampler="sobol", rfilter="box")
    a = stress_scene(4, 24, 24, 40, 40)
    b = stress_scene(4, 24, 24, 40, 40, instanced=True)
    for m in a.meshes:
        if m.name.startswith("inst"):
            m.bsdf = b.meshes[0].bsdf
    fa, sa = O.OracleScene(a).render(rp)
    ob = O.OracleScene(b)
    fb, sb = ob.render(rp)
    ra, rb = O.develop(fa), O.develop(fb)
    assert np.sqrt(((ra - rb) ** 2).sum() / (ra ** 2).sum()) < 2e-3
    assert abs(sa["pathLengthSum"] - sb["pathLengthSum"]) <= 2e-3 * sa["pathLengthSum"]
    # kd-tree and brute force agree exactly through the instances as well
    fc, _ = O.OracleScene(b, use_tree=False).render(rp)
    assert np.array_equal(fb, fc)
    assert b.n_triangles() * 3 < a.n_triangles()


def test_thinlens_sensor_focus_and_blur():
    """thinlens.cpp:327-350: a small emitter in the focal plane images sharply for any aperture; out of focus its image spreads into
    a disc of the predicted diameter; a tiny aperture reproduces the pinhole image."""
    import dataclasses
    from mitsuba_b200.scene import _quad
    P, I = _quad([(-0.05, -0.05, 5), (-0.05, 0.05, 5), (0.05, 0.05, 5), (0.05, -0.05, 5)], (0, 0, -1))
    light = Mesh(P, I, bsdf=Bsdf("diffuse", reflectance=(0, 0, 0)), radiance=(1, 1, 1))
    cam = Camera(look_at((0, 0, 0), (0, 0, 1), (0, 1, 0)), fov=20.0, near=0.1, far=100.0, width=64, height=64)
    rp = RenderParams(spp=64, sampler="independent", rfilter="box", max_depth=1)

    def lit_pixels(c):
        film, _ = O.OracleScene(SceneDesc([light], c)).render(rp)
        return int((O.develop(film).sum(2) > 0.02).sum()), float(O.develop(film).sum())
    n_pin, e_pin = lit_pixels(cam)
    n_focus, e_focus = lit_pixels(dataclasses.replace(cam, aperture_radius=0.2, focus_distance=5.0))
    n_blur, e_blur = lit_pixels(dataclasses.replace(cam, aperture_radius=0.2, focus_distance=2.5))
    n_tiny, e_tiny = lit_pixels(dataclasses.replace(cam, aperture_radius=1e-5, focus_distance=2.5))
    # quad: 0.1 wide at z = 5 -> 0.1 / (2 * 5 * tan(10 deg)) * 64 = 3.6 pixels across
    assert 9 <= n_pin <= 25 and abs(n_focus - n_pin) <= 6 and abs(n_tiny - n_pin) <= 2
    # defocus: circle of confusion at the focal plane z = 2.5 has radius 0.2 * (5 - 2.5) / 5 = 0.1 -> 0.2 / (2 * 2.5 * tan(10 deg)) * 64 = 14.5 px
    # across (plus the quad): > 100 lit pixels
    assert n_blur > 100
    # the lens collects the same flux in all cases (energy spreads, it is not lost)
    assert abs(e_focus - e_pin) < 0.1 * e_pin and abs(e_blur - e_pin) < 0.1 * e_pin
