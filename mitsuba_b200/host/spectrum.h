// Spectral power distributions -> linear RGB on the host, for <spectrum filename="x.spd"> and <spectrum value="l0:v0, l1:v1, ..."> of the
// scene file (scenehandler.cpp:557-611: InterpolatedSpectrum, zeroExtend, Spectrum::fromContinuousSpectrum, clampNegative).  Scene set-up,
// not on the render path.
#pragma once
#include <string>
#include <vector>

namespace b2host {

// (wavelength nm, value) samples in increasing wavelength -> ITU-R BT.709 linear RGB.  X, Y, Z = the integral of the piecewise-linear spectrum
// times the CIE 1931 matching functions over 360..830 nm, divided by the integral of ybar (spectrum.cpp:172-186); the integrals are evaluated
// exactly (product of two piecewise-linear functions, float64), where the reference runs an adaptive Gauss-Lobatto quadrature with a 1e-4
// tolerance.  zeroExtend: add a zero sample one average spacing before / after the data when the end values are not zero (spectrum.cpp:630-648).
// Negative components are clamped to zero.  Returns false with `err` set when the table is missing or the samples are unusable.
bool spectrumToRGB(std::vector<double> wavelengths, std::vector<double> values, bool zeroExtend, float rgb[3], std::string &err);

// "lambda value" lines, '#' comments (InterpolatedSpectrum(path), spectrum.cpp:575-602)
bool readSpd(const std::string &path, std::vector<double> &wavelengths, std::vector<double> &values, std::string &err);

} // namespace b2host
