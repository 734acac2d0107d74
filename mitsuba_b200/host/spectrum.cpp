#include "spectrum.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>

extern "C" const char *b2_data_dir_(void); // b2_host.cpp: <dir of libb2mts.so>/data or $B2MTS_DATA

namespace b2host {
namespace {

struct Observer { std::vector<double> l, x, y, z; bool ok = false; std::string err; };

const Observer &observer() {
    static Observer o;
    static std::once_flag once;
    std::call_once(once, []() {
        const std::string path = std::string(b2_data_dir_()) + "/cie1931_xyz_1nm.txt";
        std::ifstream f(path);
        if (!f) { o.err = "cannot open \"" + path + "\""; return; }
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream is(line);
            std::string t[4];
            if (!(is >> t[0] >> t[1] >> t[2] >> t[3])) continue;
            o.l.push_back(strtod(t[0].c_str(), nullptr)); o.x.push_back(strtod(t[1].c_str(), nullptr)); // hex floats
            o.y.push_back(strtod(t[2].c_str(), nullptr)); o.z.push_back(strtod(t[3].c_str(), nullptr));
        }
        o.ok = o.l.size() >= 2;
        if (!o.ok) o.err = "\"" + path + "\" holds no table";
    });
    return o;
}

// piecewise-linear f over knots xs (zero outside)
double evalLinear(const std::vector<double> &xs, const std::vector<double> &fs, double x) {
    if (x < xs.front() || x > xs.back()) return 0.0;
    const size_t i = (size_t) (std::upper_bound(xs.begin(), xs.end(), x) - xs.begin());
    if (i == 0) return fs.front();
    if (i >= xs.size()) return fs.back();
    const double t = (x - xs[i - 1]) / (xs[i] - xs[i - 1]);
    return fs[i - 1] + t * (fs[i] - fs[i - 1]);
}

// exact integral over [lo, hi] of f * g, both piecewise linear inside their own knot range and zero outside it: the integration runs over the
// intersection of the ranges (so the jumps at the ends of the data lie on the interval ends), and on a segment without interior knots the
// product is a quadratic whose integral is h * ((fa * ga + fb * gb) / 3 + (fa * gb + fb * ga) / 6)
double integrateProduct(const std::vector<double> &xf, const std::vector<double> &f, const std::vector<double> &xg, const std::vector<double> &g, double lo, double hi) {
    lo = std::max(lo, std::max(xf.front(), xg.front()));
    hi = std::min(hi, std::min(xf.back(), xg.back()));
    if (!(hi > lo)) return 0.0;
    std::vector<double> knots;
    knots.push_back(lo);
    for (double k : xf) if (k > lo && k < hi) knots.push_back(k);
    for (double k : xg) if (k > lo && k < hi) knots.push_back(k);
    knots.push_back(hi);
    std::sort(knots.begin(), knots.end());
    double sum = 0;
    double fa = evalLinear(xf, f, lo), ga = evalLinear(xg, g, lo);
    for (size_t i = 0; i + 1 < knots.size(); ++i) {
        const double a = knots[i], b = knots[i + 1], h = b - a;
        const double fb = evalLinear(xf, f, b), gb = evalLinear(xg, g, b);
        if (h > 0) sum += h * ((fa * ga + fb * gb) / 3.0 + (fa * gb + fb * ga) / 6.0);
        fa = fb; ga = gb;
    }
    return sum;
}

} // namespace

bool spectrumToRGB(std::vector<double> wl, std::vector<double> val, bool zeroExtend, float rgb[3], std::string &err) {
    const Observer &o = observer();
    if (!o.ok) { err = "CIE observer table: " + o.err; return false; }
    if (wl.size() != val.size() || wl.size() < 2) { err = "a spectrum needs at least 2 (wavelength, value) entries"; return false; }
    for (size_t i = 1; i < wl.size(); ++i)
        if (!(wl[i] > wl[i - 1])) { err = "InterpolatedSpectrum: spectral power distribution values must be provided in order of increasing wavelength!"; return false; } // spectrum.cpp:617-620
    if (zeroExtend) { // spectrum.cpp:630-648 (float arithmetic there; the knot position moves the result far below the tolerance of the comparison)
        const double spacing = (wl.back() - wl.front()) / (double) (wl.size() - 1);
        if (val.front() != 0) { wl.insert(wl.begin(), wl.front() - spacing); val.insert(val.begin(), 0.0); }
        if (val.back() != 0) { wl.push_back(wl.back() + spacing); val.push_back(0.0); }
    }
    const double lo = o.l.front(), hi = o.l.back();
    std::vector<double> one(o.l.size(), 1.0);
    const double X = integrateProduct(wl, val, o.l, o.x, lo, hi), Y = integrateProduct(wl, val, o.l, o.y, lo, hi), Z = integrateProduct(wl, val, o.l, o.z, lo, hi);
    const double norm = 1.0 / integrateProduct(o.l, one, o.l, o.y, lo, hi);
    const double x = X * norm, y = Y * norm, z = Z * norm;
    // XYZ -> ITU-R BT.709 linear RGB (spectrum.cpp:222-227)
    const double r = 3.240479 * x + -1.537150 * y + -0.498535 * z, g = -0.969256 * x + 1.875991 * y + 0.041556 * z, b = 0.055648 * x + -0.204043 * y + 1.057311 * z;
    rgb[0] = (float) std::max(r, 0.0); rgb[1] = (float) std::max(g, 0.0); rgb[2] = (float) std::max(b, 0.0);
    return true;
}

bool readSpd(const std::string &path, std::vector<double> &wl, std::vector<double> &val, std::string &err) {
    std::ifstream f(path);
    if (!f) { err = "InterpolatedSpectrum: could not open \"" + path + "\""; return false; }
    std::string line;
    wl.clear(); val.clear();
    while (std::getline(f, line)) {
        const size_t a = line.find_first_not_of(" \t\r\n");
        if (a == std::string::npos || line[a] == '#') continue;
        std::istringstream is(line.substr(a));
        double l, v;
        if (!(is >> l >> v)) break;
        wl.push_back(l); val.push_back(v);
    }
    if (wl.empty()) { err = "\"" + path + "\": unable to parse any entries!"; return false; }
    return true;
}

} // namespace b2host

// Host-only C entry (no device): the conversion above for n samples; 0, or -1 with the message in err
extern "C" int b2_spectrum_to_rgb(const float *wavelengths, const float *values, int n, int zero_extend, float rgb[3], char *err, int errLen) {
    std::string e;
    std::vector<double> wl(wavelengths, wavelengths + std::max(0, n)), val(values, values + std::max(0, n));
    if (b2host::spectrumToRGB(wl, val, zero_extend != 0, rgb, e)) return 0;
    if (err && errLen > 0) { strncpy(err, e.c_str(), (size_t) errLen - 1); err[errLen - 1] = 0; }
    return -1;
}
