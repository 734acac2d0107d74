/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
 * Samplers: sobol (src/samplers/sobol.cpp + sobolseq.h), independent over SFMT19937
 * (src/samplers/independent.cpp + src/libcore/random.cpp), and the counter-based TEA stream that
 * the GPU uses in place of the thread-order-dependent SFMT stream (documented deviation,
 * SURVEY.md 0.5). */
#pragma once
#include "orc_math.h"
#include <vector>

namespace orc {

/* ---- TEA: include/mitsuba/core/qmc.h:146-156 ---- */
inline uint64_t sampleTEA(uint32_t v0, uint32_t v1, int rounds = 4) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9;
        v0 += ((v1 << 4) + 0xA341316C) ^ (v1 + sum) ^ ((v1 >> 5) + 0xC8013EA4);
        v1 += ((v0 << 4) + 0xAD90777D) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7E95761E);
    }
    return ((uint64_t) v1 << 32) + v0;
}

/* ---- Sobol tables (data extracted by tools/extract_sobol_tables.py; see that file) ---- */
struct SobolTables {
    const uint32_t *m32 = nullptr;   /* [1024][52]  sobolseq.h:31-38 */
    const uint64_t *vdc = nullptr;   /* [25][52]    vdc_sobol_matrices     */
    const uint64_t *inv = nullptr;   /* [26][52]    vdc_sobol_matrices_inv */
};
static const uint32_t kSobolDims = 1024, kSobolSize = 52;

/* src/samplers/sobolseq.h:45-60 sampleSingle */
inline float sobolSample(const SobolTables &T, uint64_t index, uint32_t dimension, uint32_t scramble) {
    uint32_t result = scramble;
    for (uint32_t i = dimension * kSobolSize; index; index >>= 1, ++i)
        if (index & 1) result ^= T.m32[i];
    return std::min(result * (1.0f / (1ULL << 32)), kOneMinusEps);
}
/* src/samplers/sobolseq.h:104-133 look_up (SINGLE_PRECISION scramble branch) */
inline uint64_t sobolLookUp(const SobolTables &T, uint32_t m, uint32_t frame, uint32_t px, uint32_t py,
                            uint64_t scramble) {
    const uint32_t m2 = m << 1;
    uint64_t index = uint64_t(frame) << m2;
    uint64_t delta = 0;
    for (uint32_t c = 0; frame; frame >>= 1, ++c)
        if (frame & 1) delta ^= T.vdc[(m - 1) * kSobolSize + c];
    scramble = (scramble & 0xFFFFFFFF) >> (32 - m);
    uint64_t b = (((uint64_t) (px ^ scramble) << m) | (py ^ scramble)) ^ delta;
    for (uint32_t c = 0; b; b >>= 1, ++c)
        if (b & 1) index ^= T.inv[(m - 1) * kSobolSize + c];
    return index;
}

/* math::roundToPowerOfTwo / log2i (include/mitsuba/core/math.h) */
inline uint32_t roundToPowerOfTwo(uint32_t i) {
    i--; i |= i >> 1; i |= i >> 2; i |= i >> 4; i |= i >> 8; i |= i >> 16;
    return i + 1;
}
inline uint32_t log2i(uint32_t v) { uint32_t r = 0; while (v >>= 1) ++r; return r; }

/* ---- SFMT19937: src/libcore/random.cpp:66-96 (parameters), 118-147 (shifts), 181-195
 * (do_recursion), 288-296 (gen_rand64), 318-343 (period_certification), 372-393 (gen_rand_all,
 * scalar branch), 397-405 (init_gen_rand), 630-640 (nextFloat) ---- */
struct SFMT {
    enum { MEXP = 19937, N = MEXP / 128 + 1, N32 = N * 4, N64 = N * 2, POS1 = 122, SL1 = 18, SL2 = 1, SR1 = 11, SR2 = 1 };
    union W128 { uint64_t u64[2]; uint32_t u[4]; };
    union { W128 sfmt[N]; uint32_t p32[N32]; uint64_t p64[N64]; };
    int idx;
    static void rshift128(W128 &out, const W128 &in, int shift) {
        uint64_t th = in.u64[1], tl = in.u64[0];
        uint64_t oh = th >> (shift * 8), ol = tl >> (shift * 8);
        ol |= th << (64 - shift * 8);
        out.u64[0] = ol; out.u64[1] = oh;
    }
    static void lshift128(W128 &out, const W128 &in, int shift) {
        uint64_t th = in.u64[1], tl = in.u64[0];
        uint64_t oh = th << (shift * 8), ol = tl << (shift * 8);
        oh |= tl >> (64 - shift * 8);
        out.u64[0] = ol; out.u64[1] = oh;
    }
    static void recursion(W128 &r, const W128 &a, const W128 &b, const W128 &c, const W128 &d) {
        static const uint32_t MSK[4] = {0xdfffffefU, 0xddfecb7fU, 0xbffaffffU, 0xbffffff6U};
        W128 x, y;
        lshift128(x, a, SL2);
        rshift128(y, c, SR2);
        W128 out;
        for (int i = 0; i < 4; ++i)
            out.u[i] = a.u[i] ^ x.u[i] ^ ((b.u[i] >> SR1) & MSK[i]) ^ y.u[i] ^ (d.u[i] << SL1);
        r = out;
    }
    void periodCertification() {
        static const uint32_t parity[4] = {0x00000001U, 0x00000000U, 0x00000000U, 0x13c9e684U};
        int inner = 0;
        for (int i = 0; i < 4; ++i) inner ^= p32[i] & parity[i];
        for (int i = 16; i > 0; i >>= 1) inner ^= inner >> i;
        inner &= 1;
        if (inner == 1) return;
        for (int i = 0; i < 4; ++i) {
            uint32_t work = 1;
            for (int j = 0; j < 32; ++j) {
                if ((work & parity[i]) != 0) { p32[i] ^= work; return; }
                work <<= 1;
            }
        }
    }
    void seed(uint64_t s) {
        p64[0] = s;
        for (int i = 1; i < N64; ++i)
            p64[i] = (6364136223846793005ULL * (p64[i - 1] ^ (p64[i - 1] >> 62)) + i);
        idx = N32;
        periodCertification();
    }
    void genAll() {
        W128 *r1 = &sfmt[N - 2], *r2 = &sfmt[N - 1];
        int i;
        for (i = 0; i < N - POS1; ++i) {
            recursion(sfmt[i], sfmt[i], sfmt[i + POS1], *r1, *r2);
            r1 = r2; r2 = &sfmt[i];
        }
        for (; i < N; ++i) {
            recursion(sfmt[i], sfmt[i], sfmt[i + POS1 - N], *r1, *r2);
            r1 = r2; r2 = &sfmt[i];
        }
    }
    uint64_t nextULong() {
        if (idx >= N32) { genAll(); idx = 0; }
        uint64_t r = p64[idx / 2];
        idx += 2;
        return r;
    }
    float nextFloat() {
        union { uint32_t u; float f; } x;
        x.u = (uint32_t) ((nextULong() & 0xFFFFFFFF) >> 9) | 0x3f800000UL;
        return x.f - 1.0f;
    }
};

/* include/mitsuba/render/sampler.h:77-117 (the subset `path` uses) */
struct Sampler {
    virtual ~Sampler() {}
    virtual void generate(int px, int py) = 0;           /* per pixel */
    virtual void advance() = 0;                          /* next sample of the pixel */
    virtual float next1D() = 0;
    virtual void next2D(float &a, float &b) = 0;
    /* true once the current sample ran past the sampler's dimension table (Sobol': the reference aborts the render,
       sobol.cpp:223-225; volpath here ends the path instead) */
    virtual bool exhausted() const { return false; }
};

/* src/samplers/sobol.cpp:147-158,167-216,218-252.  No sample arrays are requested by `path`, so m_arrayStartDim ==
 * m_arrayEndDim == 5 (:105,:170-172).  next1D's skip test (:220-221, `dim >= start && dim < end`) then never fires, but next2D's
 * (:231-232, `dim + 1 >= start && dim < end`) does for dim == 4: a 2-D request that would straddle the (empty) array range jumps to
 * dimension 5, i.e. Sobol' dimension 4 is never used by a sequence of next2D() calls.  Found by rendering with the reference's own
 * sampler + integrator sources (oracle/path_ref_shim.cpp); restated here and in the device sampler. */
struct SobolSampler : Sampler {
    const SobolTables *T;
    uint64_t scramble = 0;      /* after the TEA step of sobol.cpp:96-102 */
    float resolution = 1;
    uint32_t logResolution = 0;
    int px = 0, py = 0;
    uint64_t sampleIndex = 0, sobolIndex = 0;
    uint32_t dimension = 0;
    bool dimOverflow = false, sampleOverflow = false;
    SobolSampler(const SobolTables *t, uint64_t scrambleProp, int filmW, int filmH) : T(t) {
        if (scrambleProp) scramble = sampleTEA((uint32_t) scrambleProp, (uint32_t) (scrambleProp >> 32));
        /* setFilmResolution(res, bucketed=true) -- integrator.cpp:38-42 */
        uint32_t res = roundToPowerOfTwo((uint32_t) std::max(filmW, filmH));
        resolution = (float) res;
        logResolution = log2i(res);
    }
    bool exhausted() const override { return sampleOverflow; }
    void setSampleIndex(uint64_t i) {
        dimension = 0; sampleOverflow = false;
        sampleIndex = i;
        if (logResolution > 1 && px >= 0)
            sobolIndex = sobolLookUp(*T, logResolution, (uint32_t) sampleIndex, (uint32_t) px, (uint32_t) py, scramble);
        else
            sobolIndex = sampleIndex;
    }
    void generate(int x, int y) override { px = x; py = y; setSampleIndex(0); }
    void advance() override { setSampleIndex(sampleIndex + 1); }
    float next1D() override {
        if (dimension >= kSobolDims) { dimOverflow = sampleOverflow = true; dimension = kSobolDims - 1; } /* reference: Log(EError) */
        return sobolSample(*T, sobolIndex, dimension++, (uint32_t) scramble);
    }
    void next2D(float &a, float &b) override {
        if (dimension + 1 >= 5 && dimension < 5) dimension = 5; /* sobol.cpp:231-232 with m_arrayStartDim == m_arrayEndDim == 5 */
        if (dimension + 1 >= kSobolDims) { dimOverflow = sampleOverflow = true; dimension = kSobolDims - 2; }
        if (dimension == 0 && sobolIndex != sampleIndex) {
            a = sobolSample(*T, sobolIndex, dimension++, (uint32_t) scramble) * resolution - px;
            b = sobolSample(*T, sobolIndex, dimension++, (uint32_t) scramble) * resolution - py;
        } else {
            a = sobolSample(*T, sobolIndex, dimension++, (uint32_t) scramble);
            b = sobolSample(*T, sobolIndex, dimension++, (uint32_t) scramble);
        }
    }
};

/* src/samplers/independent.cpp:82-103: a sequential SFMT stream (one per worker thread in the
 * reference, renderjob.cpp:58-65; here one per oracle thread, seeded seed+thread). */
struct IndependentSampler : Sampler {
    SFMT rng;
    explicit IndependentSampler(uint64_t seed) { rng.seed(seed); }
    void generate(int, int) override {}
    void advance() override {}
    float next1D() override { return rng.nextFloat(); }
    void next2D(float &a, float &b) override { a = rng.nextFloat(); b = rng.nextFloat(); }
};

/* Counter-based stream keyed by (pixel linear index, sample index, dimension): the GPU's
 * `independent` (DESIGN.md "samplers"); restated here so the GPU path can be checked sample for
 * sample.  (lo, hi) = TEA8(v0 = pixel*spp + sample (mod 2^32) ^ seed_lo, v1 = (dim >> 1) ^ seed_hi); even dim -> lo, odd dim -> hi;
 * float via the MTGP trick of random.cpp:630-640. */
struct CounterSampler : Sampler {
    uint32_t W, spp, seedLo, seedHi, key = 0, dim = 0, s = 0;
    int px = 0, py = 0;
    CounterSampler(int w, uint32_t spp_, uint64_t seed) : W((uint32_t) w), spp(spp_), seedLo((uint32_t) seed), seedHi((uint32_t) (seed >> 32)) {}
    void rekey() { key = (((uint32_t) py * W + (uint32_t) px) * spp + s) ^ seedLo; dim = 0; }
    void generate(int x, int y) override { px = x; py = y; s = 0; rekey(); }
    void advance() override { ++s; rekey(); }
    float next1D() override {
        /* one 8-round TEA block (with 4 rounds consecutive keys give correlated words: chi^2 fails) serves two dimensions:
           2k -> low word, 2k+1 -> high word */
        uint32_t d = dim++;
        uint64_t r = sampleTEA(key, (d >> 1) ^ seedHi, 8);
        uint32_t w = (d & 1u) ? (uint32_t) (r >> 32) : (uint32_t) r;
        union { uint32_t u; float f; } x;
        x.u = (w >> 9) | 0x3f800000UL;
        return x.f - 1.0f;
    }
    void next2D(float &a, float &b) override { a = next1D(); b = next1D(); }
};

} // namespace orc
