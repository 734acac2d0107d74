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


def _srgb_to_linear(v):
    v = v.astype(np.float32)
    return np.where(v <= np.float32(0.04045), v * np.float32(1 / 12.92), np.power((v + np.float32(0.055)) * np.float32(1 / 1.055), np.float32(2.4))).astype(np.float32)


def test_png_files_decode_like_libpng_and_become_linear_like_the_reference_bitmap():
    """PNG (bitmap.cpp:2465-2555 through libpng; here an own reader on zlib): files written by libpng (8 / 16 bit, grey / RGB / RGBA, one that
    uses every scan-line filter) and hand-assembled ones (4-bit palette, 2-bit grey, a gAMA chunk; filter types 0-4 in turn) decode to the same
    integer samples; samples / 255 (65535) go through the sRGB curve -- or through the exponent 1 / gAMA -- as Bitmap::convert does
    (fmtconv.cpp:1092-1101); alpha is dropped (bitmap.cpp:270-275)."""
    want = np.load(os.path.join(IMAGES, "png_expected.npz"))
    assert len(want.files) == 9
    for name in want.files:
        ints = want[name]
        v = ints.astype(np.float32) * (np.float32(1 / 65535) if ints.dtype == np.uint16 else np.float32(1 / 255))
        lin = np.power(v, np.float32(np.float32(1.0) / np.float32(0.45455))) if name == "rgb8_gama.png" else _srgb_to_linear(v)
        got = api.load_image(os.path.join(IMAGES, name))
        assert got.shape == lin.shape, (name, got.shape, lin.shape)
        assert np.allclose(got, lin, rtol=2e-6, atol=1e-7), (name, float(np.abs(got - lin).max()))
        raw = api.load_image(os.path.join(IMAGES, name), gamma=1.0)                     # a `gamma` property of 1: the samples themselves
        assert np.array_equal(raw, v), name


def test_pfm_and_ppm_files(tmp_path):
    """The two formats the loader read first (Bitmap::readPFM / readPPM, bitmap.cpp:3764-3814, :3857-3895): PFM is linear and stored bottom row
    first, little- or big-endian by the sign of its scale; binary 8-bit PPM goes through the sRGB curve."""
    rng = np.random.default_rng(2)
    img = rng.random((5, 7, 3)).astype(np.float32)
    (tmp_path / "le.pfm").write_bytes(b"PF\n7 5\n-1.0\n" + img[::-1].astype("<f4").tobytes())
    (tmp_path / "be.pfm").write_bytes(b"PF\n7 5\n2.0\n" + img[::-1].astype(">f4").tobytes())
    (tmp_path / "grey.pfm").write_bytes(b"Pf\n7 5\n-1.0\n" + img[::-1, :, 0].astype("<f4").tobytes())
    assert np.array_equal(api.load_image(tmp_path / "le.pfm"), img)
    assert np.array_equal(api.load_image(tmp_path / "be.pfm"), img * np.float32(2.0))
    assert np.array_equal(api.load_image(tmp_path / "grey.pfm")[..., 0], img[..., 0])
    raw = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    (tmp_path / "a.ppm").write_bytes(b"P6\n7 5\n255\n" + raw.tobytes())
    assert np.allclose(api.load_image(tmp_path / "a.ppm"), _srgb_to_linear(raw.astype(np.float32) * np.float32(1 / 255)), rtol=2e-6, atol=1e-7)
    (tmp_path / "x.bin").write_bytes(b"GIF89a....")
    with pytest.raises(api.B2Error, match="unsupported image format"):
        api.load_image(tmp_path / "x.bin")
