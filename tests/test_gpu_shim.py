"""The drop-in boundary, compiled: reference Scene objects (built by the reference's own classes inside oracle/_ref/libb200shim.so, exactly
as for the oracle pins) rendered through the Mitsuba-side plugin of this repository -- class B200PathTracer : public Integrator,
mitsuba_b200/host/b200_integrator.cpp, which marshals the Scene into the C-ABI of libb2mts.so and hands the film back through
Film::setBitmap -- and compared with the films the reference's own MIPathTracer + renderBlock produced for the same scenes
(tests/golden/path_ref*.npz).  No oracle and no Python scene marshalling in between: Scene -> plugin -> C-ABI -> CUDA -> film."""
import ctypes as C
import os

import numpy as np
import pytest

import ref_pins

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "..", "oracle", "_ref", "libb200shim.so")

# cases the 0.6 object API lets the plugin marshal: `path`, Sobol' sampler, BSDFs without children, no media, no instances
SUPPORTED = ["cbox_box_8spp", "cbox_gaussian_16spp", "cbox_depth3_scramble", "cbox_strict_hidden", "ball_roughconductor_ggx",
             "ball_roughdielectric_beckmann", "ball_dielectric", "ball_conductor", "ball_roughconductor_as", "ball_plastic",
             "thinlens_cbox_sobol", "env_only_ball", "env_plus_area_cbox", "crop_cbox_sobol",
             "envmap_only_ball", "envmap_hidden_cbox", "envmap_glass_ball"]   # EnvironmentMap: image, scale, toWorld read back through the Emitter interface


def rel_l2(a, b):
    return float(np.sqrt(((a.astype(np.float64) - b) ** 2).sum() / (b.astype(np.float64) ** 2).sum()))


@pytest.fixture(scope="module")
def shim(b2ctx):   # b2ctx: the session's device check (conftest.py)
    if not os.path.exists(SHIM):
        pytest.fail("oracle/_ref/libb200shim.so is missing: run __graft_entry__.build() where the reference tree is present")
    lib = C.CDLL(SHIM)
    lib.pathref_render_b200.restype = C.c_int
    return lib


def render_through_plugin(lib, desc, rp, parity=True):
    h = ref_pins.reference_scene(lib, desc, rp)
    W, H = desc.camera.film_size()
    out = np.zeros((H, W, 5), np.float32)
    err = C.create_string_buffer(2048)
    rc = lib.pathref_render_b200(h, 0, int(parity), out.ctypes.data_as(C.POINTER(C.c_float)), err, 2048)
    return rc, out, err.value.decode(errors="replace")


def test_reference_scenes_render_through_the_mitsuba_side_plugin(shim):
    g = {**np.load(os.path.join(HERE, "golden", "path_ref.npz")), **np.load(os.path.join(HERE, "golden", "path_ref_ext.npz")),
         **np.load(os.path.join(HERE, "golden", "path_ref_env.npz"))}
    cases = {name: (desc, rp) for name, desc, rp in list(ref_pins.image_cases()) + list(ref_pins.image_cases_ext()) + list(ref_pins.image_cases_env())}
    for name in SUPPORTED:
        desc, rp = cases[name]
        ref = g[name + "/film"]
        rc, film, err = render_through_plugin(shim, desc, rp)
        assert rc == 0, (name, err)
        film = film.reshape(ref.shape)
        assert np.allclose(film[..., 4], ref[..., 4], rtol=1e-5, atol=1e-6), name       # weights: identical sample positions and splats
        assert np.allclose(film[..., 3], ref[..., 3], rtol=1e-4, atol=1e-4), name       # alpha
        tol = 1e-3 if name.startswith(("crop", "envmap")) else 3e-4                       # as tests/test_gpu_z_reference_images_ext.py, test_gpu_envmap.py
        assert rel_l2(film[..., :3], ref[..., :3]) <= tol, (name, rel_l2(film[..., :3], ref[..., :3]))


def test_plugin_reports_what_the_object_api_hides(shim):
    """Nested BSDFs are private members of their parents in Mitsuba 0.6: the plugin says so instead of rendering something else."""
    cases = {name: (desc, rp) for name, desc, rp in ref_pins.image_cases()}
    for name, fragment in (("ball_coating_diffuse", "wraps another BSDF"), ("ball_twosided_two", "wraps another BSDF")):
        rc, _, err = render_through_plugin(shim, *cases[name])
        assert rc == 1 and fragment in err, (name, err)
    desc, rp = cases["vol_cbox_sobol"]     # the plugin is the `path` integrator: same scene, same film as path (no media in it)
    rc, film, err = render_through_plugin(shim, desc, rp)
    assert rc == 0, err
