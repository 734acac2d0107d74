"""The oracle's component functions against the REFERENCE's own code.

oracle/core_ref_shim.cpp compiles the reference's microfacet.h, triaccel.h, aabb.h, triangle.cpp, warp.cpp, util.cpp, quad.cpp, math.cpp,
pmf.h and qmc.h from where they lie under /root/reference (behind the stand-in headers of oracle/shim_core/) into
oracle/_ref/libcoreref.so; tests/gen_golden.py ran it on the seeded inputs of tests/ref_pins.py and committed the outputs as
tests/golden/core_ref.npz.  The oracle must reproduce every value bit for bit (NaN = NaN)."""
import ctypes as C
import os

import numpy as np
import pytest

import ref_pins
from oracle import oracle_api as O

HERE = os.path.dirname(os.path.abspath(__file__))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a, nan=0.0), np.nan_to_num(b, nan=0.0))
    return np.array_equal(a, b)


@pytest.fixture(scope="module")
def oracle_outputs():
    return ref_pins.run(O.lib(), "orc_", ref_pins.inputs())


def test_oracle_components_match_reference_golden(oracle_outputs):
    g = np.load(os.path.join(HERE, "golden", "core_ref.npz"))
    assert len(g.files) == len(oracle_outputs) == 71
    bad = [k for k in g.files if not same(g[k], oracle_outputs[k])]
    assert not bad, bad
    # the fixture is not vacuous: hits and misses, degenerate triangles, total internal reflection, zero-weight pmf bins
    assert 0.2 < g["triaccel_hits"][:, 0].mean() < 0.95 and g["triaccel_status"].sum() >= 6
    assert (g["fresnel_dielectric_0.6667"][:, 0] == 1.0).sum() > 10 and g["aabb"][:, 0].min() == 0 and g["aabb"][:, 0].max() == 1


def test_oracle_components_match_live_reference_when_present(oracle_outputs):
    so = os.path.join(HERE, "..", "oracle", "_ref", "libcoreref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libcoreref.so not built (the reference tree is not on this machine)")
    lib = C.CDLL(so)
    ref = ref_pins.run(lib, "coreref_", ref_pins.inputs())
    assert not [k for k in ref if not same(ref[k], oracle_outputs[k])]
    x = ref_pins.inputs(seed=99)  # and on inputs the fixture has not seen
    assert not [k for k, v in ref_pins.run(lib, "coreref_", x).items() if not same(v, ref_pins.run(O.lib(), "orc_", x)[k])]
    lib.coreref_fresnel_diffuse_reflectance.restype = C.c_float
    from mitsuba_b200.scene import fresnel_diffuse_reflectance
    for eta in (1.5, 1.0 / 1.5, 1.33, 1.49):  # util.cpp:807-859 with fast = false (Gauss-Lobatto): the loaders integrate it numerically
        assert abs(lib.coreref_fresnel_diffuse_reflectance(C.c_float(eta), 0) - fresnel_diffuse_reflectance(eta)) < 2e-6


# plastic evaluates util.cpp's fresnelDiffuseReflectance(1/eta, fast = false) -- an adaptive Gauss-Lobatto quadrature in float -- once
# at construction; the host side of this repository integrates the same function with Simpson's rule in double, which can differ in
# the last bit of that constant.  Everything else is bit-exact.
LAST_BIT_OF_A_CONSTANT = {"plastic": 5e-5, "twosided_two": 5e-5}


def _bsdf_close(name, ref, got):
    tol = LAST_BIT_OF_A_CONSTANT.get(name)
    for k, r in ref.items():
        g = np.asarray(got[k])
        r = np.asarray(r)
        if tol is None:
            if not same(r, g):
                return k
        else:
            if r.dtype.kind != "f":
                if not np.array_equal(r, g):
                    return k
            elif not np.allclose(g, r, rtol=tol, atol=1e-7, equal_nan=True):
                return k
    return None


def test_oracle_bsdfs_match_reference_plugins_golden():
    """tests/golden/bsdf_ref.npz: outputs of the reference's own BSDF plugin sources (see tests/gen_golden.py).  16 of the 18
    configurations -- every microfacet model, coating, dielectric, conductor, twosided -- are reproduced bit for bit."""
    from bsdf_configs import configs
    g = np.load(os.path.join(HERE, "golden", "bsdf_ref.npz"))
    x = ref_pins.bsdf_inputs()
    exact = 0
    for name, b in configs().items():
        ref = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
        assert len(ref) == 8, name
        got = ref_pins.run_bsdf_oracle(b, x)
        assert _bsdf_close(name, ref, got) is None, (name, _bsdf_close(name, ref, got))
        exact += name not in LAST_BIT_OF_A_CONSTANT
        assert np.any(ref["sample"][:, 3:6] != 0), name  # the fixture holds successful samples
    assert exact == 16


def test_oracle_bsdfs_match_live_reference_plugins_when_present():
    so = os.path.join(HERE, "..", "oracle", "_ref", "libbsdfref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libbsdfref.so not built (the reference tree is not on this machine)")
    from bsdf_configs import configs
    lib = C.CDLL(so)
    x = ref_pins.bsdf_inputs(seed=4321)  # directions the fixture has not seen
    for name, b in configs().items():
        ref = ref_pins.run_bsdf_reference(lib, b, x)
        got = ref_pins.run_bsdf_oracle(b, x)
        assert _bsdf_close(name, ref, got) is None, (name, _bsdf_close(name, ref, got))


def _oracle_scene(W, H):
    from mitsuba_b200.scene import cornell_box
    return O.OracleScene(cornell_box(W, H))


def test_oracle_partials_filters_splat_and_sobol_match_reference_golden():
    """tests/golden/render_ref.npz (see tests/gen_golden.py): Intersection::computePartials incl. its degenerate branches, the
    discretised box / gaussian filters, 300 ImageBlock::put splats per filter (block borders, clamping), and the SobolSampler plugin's
    first 24 dimensions for five (scramble, resolution, spp, pixel, sample) settings -- all bit for bit."""
    g = np.load(os.path.join(HERE, "golden", "render_ref.npz"))
    got = ref_pins.run_render(O.lib(), "orc_", ref_pins.render_inputs(), _oracle_scene)
    assert len(g.files) == len(got) == 16
    assert not [k for k in g.files if not same(g[k], got[k])]
    assert g["put1_ok"].all() and g["put1_data"][..., 4].sum() > 100
    # rejected samples (imageblock.h:136-139: negative / non-finite values): closed form, no contribution
    L = O.lib()
    pos = np.array([[15.0, 25.0]] * 4, np.float32)
    val = np.array([[-1, 0, 0, 1], [np.nan, 0, 0, 1], [np.inf, 0, 0, 1], [0, 0, 0, -0.5]], np.float32)
    data, ok = np.zeros((16, 20, 5), np.float32), np.ones(4, np.int32)
    L.orc_block_put(10, 20, 16, 12, 1, C.c_float(0.5), 4, ref_pins._f(pos), ref_pins._f(val), ref_pins._f(data), ok.ctypes.data_as(C.POINTER(C.c_int)))
    assert not ok.any() and not data.any()


def test_oracle_partials_filters_splat_and_sobol_match_live_reference_when_present():
    so = os.path.join(HERE, "..", "oracle", "_ref", "librenderref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/librenderref.so not built (the reference tree is not on this machine)")
    lib = C.CDLL(so)
    x = ref_pins.render_inputs(seed=55)
    ref = ref_pins.run_render(lib, "renderref_", x)
    got = ref_pins.run_render(O.lib(), "orc_", x, _oracle_scene)
    assert not [k for k in ref if not same(ref[k], got[k])]


def test_oracle_images_match_the_reference_renderer_golden():
    """tests/golden/path_ref.npz: films rendered by a `path` renderer assembled from the reference's own sources (MIPathTracer::Li,
    renderBlock, Scene, ShapeKDTree, sensor, emitter, Sobol' sampler, filters, ImageBlock, BSDF plugins; see tests/gen_golden.py).
    Starting from the sampleToCamera matrix of the reference's sensor (camera set-up is host work), the oracle reproduces every film
    BIT FOR BIT -- rgb, alpha and weight channels -- except the plastic ball (last bit of one constructor constant, see above)."""
    g = np.load(os.path.join(HERE, "golden", "path_ref.npz"))
    n = 0
    for name, desc, rp in ref_pins.image_cases():
        ref, s2c = g[name + "/film"], g[name + "/s2c"]
        film, _ = O.OracleScene(desc, sample_to_camera=s2c).render(rp)
        film = np.asarray(film).reshape(ref.shape)
        if name == "ball_plastic":
            assert np.sqrt(((film - ref) ** 2).sum() / (ref ** 2).sum()) < 1e-6
        else:
            assert np.array_equal(film, ref), (name, float(np.abs(film - ref).max()))
        assert ref[..., :3].max() > 0.1 and ref[..., 4].min() > 0
        # and the host-side derivation of the camera matrix agrees with the reference's to float rounding
        assert np.allclose(desc.camera.sample_to_camera().astype(np.float32), s2c, rtol=1e-6, atol=1e-7)
        n += 1
    assert n == 19


def test_oracle_images_match_the_live_reference_renderer_when_present():
    so = os.path.join(HERE, "..", "oracle", "_ref", "libpathref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libpathref.so not built (the reference tree is not on this machine)")
    from mitsuba_b200.scene import RenderParams, cornell_box
    lib = C.CDLL(so)
    desc, rp = cornell_box(56, 40), RenderParams(spp=12, sampler="sobol", rfilter="gaussian", seed=3)  # a case the fixture does not hold
    ref, s2c = ref_pins.reference_render(lib, desc, rp, want_camera=True)
    film, _ = O.OracleScene(desc, sample_to_camera=s2c).render(rp)
    assert np.array_equal(np.asarray(film).reshape(ref.shape), ref)
