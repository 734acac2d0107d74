"""The image readers of the scene-file front end (b2_load_image: host-only, no device) against files written by the OpenEXR library and
OpenCV's Radiance writer, and what those libraries read back from them (tests/golden/images, tests/gen_image_fixtures.py): bit for bit."""
import os

import numpy as np
import pytest

from mitsuba_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
IMAGES = os.path.join(HERE, "golden", "images")


def test_openexr_and_rgbe_files_decode_exactly_as_the_libraries_do():
    want = np.load(os.path.join(IMAGES, "expected.npz"))
    assert len(want.files) == 11
    for name in want.files:
        got = api.load_image(os.path.join(IMAGES, name))
        assert got.shape == want[name].shape, name
        assert np.array_equal(got, want[name]), (name, float(np.abs(got - want[name]).max()))
    # a `gamma` property re-interprets the samples (bitmap.cpp:251-252): exponent, or -1 for the sRGB curve
    lin = api.load_image(os.path.join(IMAGES, "rgb_zip_float.exr"))
    assert np.allclose(api.load_image(os.path.join(IMAGES, "rgb_zip_float.exr"), gamma=2.0), lin ** 2.0, rtol=1e-6)


def test_unsupported_files_are_refused_by_name(tmp_path):
    with pytest.raises(api.B2Error, match="PIZ compression is not supported"):
        api.load_image(os.path.join(IMAGES, "rgb_piz.exr"))
    data = open(os.path.join(IMAGES, "rgb_zip_half.exr"), "rb").read()
    p = tmp_path / "cut.exr"
    p.write_bytes(data[:len(data) // 2])
    with pytest.raises(api.B2Error, match="truncated|corrupt"):
        api.load_image(str(p))
    with pytest.raises(api.B2Error, match="cannot open"):
        api.load_image(str(tmp_path / "missing.exr"))
