// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product.
//
// Transfer curves: the oracle's statement of how the composite transfer functions of the colour
// chain are evaluated per pixel since round 2 (the product's statement is
// libplacebo_b200/csrc/curves.{h,c} + plb_math.cuh plm_curve).
//
// The reference's GLSL computes e.g. the PQ EOTF as pow -> rational -> pow
// (src/shaders/colorspace.c:662-668) with an implementation-defined pow(). The tolerance of the
// path is +-1 LSB of the 16-bit output, so the composite function f is DEFINED here as a
// piecewise cubic in the position u in [-1/2, 1/2) inside the segment that holds the argument;
// segments are 2^-M of a binade wide, found from the float's bits:
//
//     seg = bits(x) >> (23 - M)         u = float(1.m') - 1.5, m' = mantissa bits below the segment
//     f(x) = c0 + u (c1 + u (c2 + u c3))                              (three fmaf)
//
// The coefficients interpolate the real-valued function at the four Chebyshev nodes of the
// segment; the real-valued function is evaluated in binary64 by the short algorithms below
// (log2 by atanh series, exp2 by Taylor series — only + - * / fma, floor, ldexp), from the same
// decimal constants the GLSL text carries. The result is within ~1.5 float ULP of the real
// function (tests/test_curves.py, against float64 closed forms) — closer than any fp32
// pow/divide/pow chain. Arguments outside the tabulated domain (negative, NaN, beyond `hi`)
// take the exact chain built on pm_powf (pl_oracle_math.h), exactly as in round 1.
// orc_set_exact_math(1) empties the domains: everything takes the exact chain (regression mode).
#ifndef PL_ORACLE_CURVE_H
#define PL_ORACLE_CURVE_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "pl_oracle_math.h"

enum orc_curve_id {
    ORC_CV_PQ_EOTF_SDR = 0, // from_pq * EOTF(x), literal constants
    ORC_CV_PQ_EOTF,         // EOTF(x)
    ORC_CV_PQ_OETF_SDR,     // OETF(to_pq * x)
    ORC_CV_PQ_OETF_DELIN,   // OETF(x / from_pq)
    ORC_CV_PQ_OETF,         // OETF(x)
    ORC_CV_PQX_OETF,        // OETF(x) with the exact float constants (peak detection)
    ORC_CV_POW_24, ORC_CV_POW_INV24, ORC_CV_POW_22, ORC_CV_POW_INV22,
    ORC_CV_COUNT,
};

struct orc_curve {
    float *coef;    // (nseg + 1) x (c0, c1, c2, c3); entry 0 = constant below the domain
    int base, nseg, mbits;
};

// segments per binade (2^M) of each curve
static inline int orc_curve_mbits(int id)
{
    if (id == ORC_CV_PQ_EOTF_SDR || id == ORC_CV_PQ_EOTF)
        return 7;
    if (id == ORC_CV_POW_24 || id == ORC_CV_POW_22)
        return 6;
    return 4;
}

// ---- the real-valued functions, binary64 ---------------------------------------------------

// log2 x, x > 0 normal: x = m 2^e, m in (2^-1/2, 2^1/2]; ln m = 2 atanh((m-1)/(m+1)), 12 terms
static inline double orc_cv_log2(double x)
{
    uint64_t b = pm_bits(x);
    int e = (int) (b >> 52) - 1023;
    double m = pm_from_bits((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) {
        m *= 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0), z = s * s;
    double p = 1.0 / 23.0;
    for (int k = 21; k >= 3; k -= 2)
        p = fma(p, z, 1.0 / (double) k);
    const double ln_m = 2.0 * fma(s * z, p, s);
    return fma(ln_m, 1.4426950408889634, (double) e);
}

// 2^t: k = round(t), 2^(t-k) = e^y with y = (t-k) ln 2, Taylor to y^16/16!
static inline double orc_cv_exp2(double t)
{
    const double k = floor(t + 0.5), y = (t - k) * 0.6931471805599453;
    double fact = 20922789888000.0; // 16!
    double q = 1.0 / fact;
    for (int j = 15; j >= 0; j--) {
        fact /= (double) (j + 1);   // j!
        q = fma(q, y, 1.0 / fact);
    }
    return ldexp(q, (int) k);
}

static inline double orc_cv_pow(double x, double y)
{
    return x > 0.0 ? orc_cv_exp2(y * orc_cv_log2(x)) : 0.0;
}

struct orc_pq_consts { double m1, m2, c1, c2, c3, inv_m1, inv_m2; };

// (c2 - c3 p) has its root at x = 1.99: the EOTF table stops at 1.5
static inline double orc_cv_pq_eotf(const struct orc_pq_consts *q, double x, double out_scale)
{
    const double p = orc_cv_pow(x, q->inv_m2), num = p - q->c1;
    if (!(num > 0.0))
        return 0.0;
    return out_scale * orc_cv_pow(num / (q->c2 - q->c3 * p), q->inv_m1);
}

static inline double orc_cv_pq_oetf(const struct orc_pq_consts *q, double x, double in_scale)
{
    const double u = orc_cv_pow(in_scale * x, q->m1);
    return orc_cv_pow((q->c1 + q->c2 * u) / (1.0 + q->c3 * u), q->m2);
}

// ---- table construction ---------------------------------------------------------------------

struct orc_curve_spec {
    int kind;             // 0 EOTF, 1 OETF, 2 power
    struct orc_pq_consts q;
    double scale, gamma;
    int emin;             // domain starts at 2^emin
    uint32_t hi_bits;     // ... and ends below the float with these bits
};

static inline double orc_curve_real(const struct orc_curve_spec *s, double x)
{
    if (s->kind == 0)
        return orc_cv_pq_eotf(&s->q, x, s->scale);
    if (s->kind == 1)
        return orc_cv_pq_oetf(&s->q, x, s->scale);
    return orc_cv_pow(x, s->gamma);
}

static inline void orc_curve_build(struct orc_curve *cv, int id, const struct orc_curve_spec *s)
{
    const int M = orc_curve_mbits(id), shift = 23 - M;
    const int first = (s->emin + 127) << M, end = (int) (s->hi_bits >> shift);
    cv->mbits = M;
    cv->base = first - 1;
    cv->nseg = end - first;
    cv->coef = malloc((size_t) (cv->nseg + 1) * 4 * sizeof(float));
    cv->coef[0] = (float) orc_curve_real(s, 0.0);
    cv->coef[1] = cv->coef[2] = cv->coef[3] = 0.0f;
    // Chebyshev nodes of [-1/2, 1/2]: +-a, +-b with a = cos(pi/8)/2, b = cos(3 pi/8)/2.
    // even part c0 + c2 u^2 from the node sums, odd part c1 u + c3 u^3 from the differences
    const double a = 0.46193976625564337, b = 0.19134171618254489;
    const double a2 = a * a, b2 = b * b, inv = 1.0 / (a2 - b2);
    for (int i = 0; i < cv->nseg; i++) {
        const double x0 = pm_ffrom_bits((uint32_t) (first + i) << shift);
        const double x1 = pm_ffrom_bits((uint32_t) (first + i + 1) << shift);
        const double w = x1 - x0, xc = x0 + 0.5 * w;
        const double fpa = orc_curve_real(s, xc + a * w), fna = orc_curve_real(s, xc - a * w);
        const double fpb = orc_curve_real(s, xc + b * w), fnb = orc_curve_real(s, xc - b * w);
        const double ea = 0.5 * (fpa + fna), oa = 0.5 * (fpa - fna);
        const double eb = 0.5 * (fpb + fnb), ob = 0.5 * (fpb - fnb);
        const double c2 = (ea - eb) * inv, c0 = eb - c2 * b2;
        const double c3 = (oa / a - ob / b) * inv, c1 = ob / b - c3 * b2;
        float *c = cv->coef + 4 * (i + 1);
        c[0] = (float) c0; c[1] = (float) c1; c[2] = (float) c2; c[3] = (float) c3;
    }
}

// ---- per-pixel evaluation -------------------------------------------------------------------

// 1 = hit (y set), 0 = x outside the domain
static inline int orc_curve_eval(const struct orc_curve *cv, float x, float *y)
{
    const uint32_t ix = pm_fbits(x);
    int idx = (int) (ix >> (23 - cv->mbits));
    idx = (idx > cv->base ? idx : cv->base) - cv->base;
    if (idx > cv->nseg)
        return 0;
    const float u = pm_ffrom_bits(((ix << cv->mbits) & 0x007FFFFFu) | 0x3F800000u) - 1.5f;
    const float *k = cv->coef + 4 * idx;
    *y = fmaf(fmaf(fmaf(k[3], u, k[2]), u, k[1]), u, k[0]);
    return 1;
}

#endif
