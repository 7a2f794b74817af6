// Generator of the transfer-curve tables (curves.h). Host C, binary64, built only from
// + - * / fma, floor, ldexp and integer bit operations, so the tables are the same bits on every
// IEEE-754 machine. The CPU oracle carries its own restatement of this definition
// (oracle/pl_oracle_curve.h); tests/test_curves.py checks both tables bit for bit and pins them
// to a float64 closed-form evaluation.
//
// Reference formulas: PQ EOTF src/shaders/colorspace.c:662-668, PQ OETF :778-784, BT.1886
// :622-630 / :759-767, gamma curves :631-650 / :768-776, peak-detection PQ :1279-1288.
#define _GNU_SOURCE
#include "internal.h"
#include "curves.h"

#include <math.h>
#include <stdint.h>

static double d_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t d_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// log2 of a positive normal double: x = m 2^e with m in (1/sqrt2, sqrt2], s = (m-1)/(m+1),
// ln m = 2 atanh s = 2 s (1 + z/3 + z^2/5 + ... + z^11/23), z = s^2 <= 0.0295
static double cv_log2(double x)
{
    const uint64_t b = d_bits(x);
    int e = (int) (b >> 52) - 1023;
    double m = d_from_bits((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
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

// 2^t, |t| < 1000: k = nearest integer, r = t - k, Taylor series of e^(r ln 2) to degree 16
static double cv_exp2(double t)
{
    const double k = floor(t + 0.5), y = (t - k) * 0.6931471805599453;
    double q = 1.0 / 20922789888000.0; // 1/16!
    static const double inv_fact[16] = {
        1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
        1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
        1.0 / 87178291200.0, 1.0 / 1307674368000.0,
    };
    for (int j = 15; j >= 0; j--)
        q = fma(q, y, inv_fact[j]);
    return ldexp(q, (int) k);
}

static double cv_pow(double x, double y) // x >= 0, y > 0
{
    return x > 0.0 ? cv_exp2(y * cv_log2(x)) : 0.0;
}

struct cv_def {
    int kind;           // 0 PQ EOTF, 1 PQ OETF, 2 power
    double m1, m2, c1, c2, c3, inv_m1, inv_m2;
    double scale;       // EOTF: output factor; OETF: input factor
    double gamma;
};

static double cv_eval(const struct cv_def *d, double x)
{
    switch (d->kind) {
    case 0: {
        const double p = cv_pow(x, d->inv_m2);
        const double num = p - d->c1;
        if (!(num > 0.0))
            return 0.0;
        return d->scale * cv_pow(num / (d->c2 - d->c3 * p), d->inv_m1);
    }
    case 1: {
        const double u = cv_pow(d->scale * x, d->m1);
        return cv_pow((d->c1 + d->c2 * u) / (1.0 + d->c3 * u), d->m2);
    }
    default:
        return cv_pow(x, d->gamma);
    }
}

static const float PQ_M1 = 2610./4096 * 1./4, PQ_M2 = 2523./4096 * 128, PQ_C1 = 3424./4096,
                   PQ_C2 = 2413./4096 * 32, PQ_C3 = 2392./4096 * 32;

static bool cv_define(int id, const struct plb_pq_consts *q, struct cv_def *d)
{
    memset(d, 0, sizeof(*d));
    d->m1 = q->m1; d->m2 = q->m2; d->c1 = q->c1; d->c2 = q->c2; d->c3 = q->c3;
    d->inv_m1 = q->inv_m1; d->inv_m2 = q->inv_m2;
    d->scale = 1.0;
    switch (id) {
    case PLB_CV_PQ_EOTF_SDR:
        d->scale = q->from_pq;
        // fall through
    case PLB_CV_PQ_EOTF:
        d->kind = 0;
        return true;
    case PLB_CV_PQ_OETF_SDR:
        d->kind = 1; d->scale = q->to_pq;
        return true;
    case PLB_CV_PQ_OETF_DELIN:
        d->kind = 1; d->scale = q->inv_from_pq;
        return true;
    case PLB_CV_PQ_OETF:
        d->kind = 1;
        return true;
    case PLB_CV_PQX_OETF:
        d->kind = 1;
        d->m1 = PQ_M1; d->m2 = PQ_M2; d->c1 = PQ_C1; d->c2 = PQ_C2; d->c3 = PQ_C3;
        return true;
    case PLB_CV_POW_24:
        d->kind = 2; d->gamma = 2.4f;
        return true;
    case PLB_CV_POW_INV24:
        d->kind = 2; d->gamma = (float) (1.0 / 2.4);
        return true;
    case PLB_CV_POW_22:
        d->kind = 2; d->gamma = 2.2f;
        return true;
    case PLB_CV_POW_INV22:
        d->kind = 2; d->gamma = (float) (1.0 / 2.2);
        return true;
    }
    return false;
}

// Cubic through the four Chebyshev nodes u = +-a, +-b of [-1/2, 1/2] (a = cos(pi/8)/2,
// b = cos(3pi/8)/2), by even / odd parts.
bool plb_curve_generate(int id, const struct plb_pq_consts *q, struct plb_curve_host *out)
{
    struct cv_def d;
    if (!cv_define(id, q, &d))
        return false;
    const int M = PLB_CV_MBITS(id), shift = 23 - M;
    const int seg_lo = PLB_CV_BASE(id) + 1;
    out->mbits = M;
    out->base = PLB_CV_BASE(id);
    out->nseg = PLB_CV_NSEG(id);
    out->coef = calloc((size_t) PLB_CV_ENTRIES(id) * 4, sizeof(float)); // last entry: padding
    if (!out->coef)
        return false;
    const double a = 0.46193976625564337, b = 0.19134171618254489;
    const double a2 = a * a, b2 = b * b, inv = 1.0 / (a2 - b2);
    out->coef[0] = (float) cv_eval(&d, 0.0);
    out->coef[1] = out->coef[2] = out->coef[3] = 0.0f;
    for (int i = 0; i < out->nseg; i++) {
        const double x0 = f_from_bits((uint32_t) (seg_lo + i) << shift);
        const double x1 = f_from_bits((uint32_t) (seg_lo + i + 1) << shift);
        const double w = x1 - x0, xc = x0 + 0.5 * w;
        const double fpa = cv_eval(&d, xc + a * w), fma_ = cv_eval(&d, xc - a * w);
        const double fpb = cv_eval(&d, xc + b * w), fmb = cv_eval(&d, xc - b * w);
        const double E1 = 0.5 * (fpa + fma_), O1 = 0.5 * (fpa - fma_);
        const double E2 = 0.5 * (fpb + fmb), O2 = 0.5 * (fpb - fmb);
        const double c2 = (E1 - E2) * inv, c0 = E2 - c2 * b2;
        const double c3 = (O1 / a - O2 / b) * inv, c1 = O2 / b - c3 * b2;
        float *c = out->coef + 4 * (i + 1);
        c[0] = (float) c0; c[1] = (float) c1; c[2] = (float) c2; c[3] = (float) c3;
    }
    return true;
}

// Diagnostics (no GPU needed): the table of curve `id` as the kernels receive it.
// coef may be NULL to query the sizes; otherwise it receives 4 * (*nseg + 1) floats.
PL_API bool pl_cuda_curve_table(int id, int *base, int *nseg, int *mbits, float *coef)
{
    struct plb_pq_consts q;
    struct plb_curve_host h;
    plb_pq_literal(&q);
    if (!plb_curve_generate(id, &q, &h))
        return false;
    if (base) *base = h.base;
    if (nseg) *nseg = h.nseg;
    if (mbits) *mbits = h.mbits;
    if (coef)
        memcpy(coef, h.coef, (size_t) (h.nseg + 1) * 4 * sizeof(float)); // without the padding entry
    free(h.coef);
    return true;
}
