"""<spectrum filename="x.spd"> / <spectrum value="l0:v0, ..."> -> linear RGB (mitsuba_b200/host/spectrum.cpp, host-only b2_spectrum_to_rgb) against
the reference's InterpolatedSpectrum + Spectrum::fromContinuousSpectrum (tests/golden/spectrum_ref.npz, tests/gen_golden.py).  The product
integrates the piecewise-linear products exactly; the reference runs an adaptive Gauss-Lobatto quadrature with a 1e-4 tolerance: they must
agree to that."""
import os

import numpy as np
import pytest

from gen_golden import spectrum_inputs
from mitsuba_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))


def test_spectra_convert_like_the_reference_loader():
    ref = np.load(os.path.join(HERE, "golden", "spectrum_ref.npz"))["rgb"]
    worst = 0.0
    for (w, v, ze), want in zip(spectrum_inputs(), ref):
        got = api.spectrum_to_rgb(w, v, zero_extend=bool(ze))
        worst = max(worst, float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)))
    assert worst < 3e-4, worst


def test_observer_table_is_the_one_the_reference_integrates_against():
    cie = np.load(os.path.join(HERE, "golden", "spectrum_ref.npz"))["cie"]
    path = os.path.join(HERE, "..", "mitsuba_b200", "data", "cie1931_xyz_1nm.txt")
    rows = np.array([[float.fromhex(t) for t in line.split()] for line in open(path) if line.strip() and not line.startswith("#")], np.float32)
    assert rows.shape == (471, 4) and np.array_equal(rows, cie)
    assert rows[0, 0] == 360 and rows[-1, 0] == 830 and abs(rows[195, 2] - 1.0) < 1e-6   # ybar peaks at 555 nm


def test_closed_forms_and_errors():
    # a flat spectrum over the whole visible range is white of that level: X = Y = Z integrals of the observer -> RGB (1, 1, 1) x level, to
    # the accuracy of the BT.709 matrix constants and the equal-energy white point (E, not D65: the channels differ by a few percent)
    rgb = api.spectrum_to_rgb([300, 900], [0.5, 0.5], zero_extend=False)
    assert abs(0.212671 * rgb[0] + 0.715160 * rgb[1] + 0.072169 * rgb[2] - 0.5) < 2e-3   # luminance = Y = the level
    # linearity and additivity
    a = api.spectrum_to_rgb([400, 500, 600, 700], [0.1, 0.4, 0.3, 0.2])
    b = api.spectrum_to_rgb([400, 500, 600, 700], [0.2, 0.8, 0.6, 0.4])
    assert np.allclose(b, 2 * a, rtol=1e-5, atol=1e-7)
    # a narrow band at 460 nm is blue, at 620 nm red (negative components clamped to zero)
    blue = api.spectrum_to_rgb([455, 460, 465], [0, 1, 0]); red = api.spectrum_to_rgb([615, 620, 625], [0, 1, 0])
    assert blue[2] > blue[1] and blue[2] > blue[0] and red[0] > red[1] and red[0] > red[2] and blue.min() >= 0 and red.min() >= 0
    # outside the observer's range nothing is seen
    assert np.array_equal(api.spectrum_to_rgb([900, 950], [1, 1], zero_extend=False), np.zeros(3, np.float32))
    with pytest.raises(api.B2Error, match="increasing wavelength"):
        api.spectrum_to_rgb([500, 400], [1, 1])
    with pytest.raises(api.B2Error, match="at least 2"):
        api.spectrum_to_rgb([500], [1])
