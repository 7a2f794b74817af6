// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product.
//
// Portable transcendental functions for the per-pixel (GLSL) arithmetic.
//
// GLSL leaves pow/exp/log/sin/cos/atan implementation-defined (every GPU vendor rounds them
// differently), so a CPU restatement has to pick one definition. libm is a poor choice: glibc,
// CUDA libdevice and a GPU driver's pow() all disagree in the last bits, and the PQ EOTF
// amplifies one float ULP of pow() error ~500x (cancellation in num/den, then ^6.28).
// Instead the oracle DEFINES these functions by a short algorithm that only uses IEEE-754
// binary64 +, *, /, fma and integer operations. Such an algorithm gives the same bits on any
// conforming machine; the CUDA kernels restate the same definition (csrc/plb_math.cuh), which
// is what makes bit-exact parity of the colour chain testable. The result is the correctly
// rounded float except when the exact value lies within ~1e-13 (relative) of a rounding
// boundary; tests/test_oracle_math.py pins every function to within 1 float ULP of glibc over
// dense sweeps.
//
// Definitions (x, y float; all intermediate arithmetic binary64, Horner steps are fma):
//   log2(x):  x = m * 2^e with m in [1, 2);  i = top 7 mantissa bits;  c_i = RCP[i] ~ 1/m with
//             9 fractional bits (so r = m c_i - 1 is exact, |r| < 2^-7.7);
//             log2(x) = e + LOGC[i] + r * A6(r),   A6 = sum_{k<=6} (-1)^(k+1) / (k ln 2) r^(k-1)
//   exp2(t):  |t| >= 160 saturates to +-160;  k = nearest integer to 64 t;  r = t - k/64;
//             exp2(t) = 2^(k >> 6) * EXP2[k & 63] * B5(r),   B5 = sum_{j<=5} (ln 2)^j / j! r^j
//             (tables RCP, LOGC = -log2(c_i), EXP2 = 2^(j/64): pl_oracle_math_tables.h, generated
//             by tools/gen_math_tables.py as exact binary64 literals)
//   pow(x,y) = float(exp2(y * log2(x)))    for finite x > 0
//   exp(x)   = float(exp2(x * log2(e))),   log(x) = float(log2(x) * ln 2)
//   sin/cos:  k = nearest integer to x * 2/pi;  r = (x - k*pio2_hi) - k*pio2_lo;
//             odd/even Taylor polynomials of degree 17/18 in r, quadrant selected by k mod 4
//   atan2(y,x): a = |y|, b = |x|, lo = min, hi = max; if lo > tan(pi/8)*hi use
//             z = (lo-hi)/(lo+hi) and offset pi/4, else z = lo/hi; atan z = z * A18(z^2);
//             then the usual octant / sign fix-ups.
#ifndef PL_ORACLE_MATH_H
#define PL_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "pl_oracle_math_tables.h"

static inline double pm_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t pm_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline uint32_t pm_fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float pm_ffrom_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// log2 of a finite, strictly positive float
static inline double pm_log2_core(float x)
{
    uint32_t ix = pm_fbits(x);
    int e = -127;
    if (ix < 0x00800000u) {          // subnormal: scale by 2^24 (exact)
        ix = pm_fbits(x * 16777216.0f);
        e -= 24;
    }
    e += (int) (ix >> 23);
    const int i = (ix >> 16) & 127;  // top 7 mantissa bits
    const double m = (double) pm_ffrom_bits((ix & 0x007FFFFFu) | 0x3F800000u); // [1, 2)
    const double r = fma(m, PM_RCP[i], -1.0);                                  // exact, |r| < 2^-7.7
    // log2(1 + r) = r * (a1 + r (a2 + ... r a6)),  a_k = (-1)^(k+1) / (k ln 2)
    double p = -0x1.ec709dc3a03fdp-3;   // -1/(6 ln 2)
    p = fma(p, r, 0x1.2776c50ef9bfep-2);  //  1/(5 ln 2)
    p = fma(p, r, -0x1.71547652b82fep-2); // -1/(4 ln 2)
    p = fma(p, r, 0x1.ec709dc3a03fdp-2);  //  1/(3 ln 2)
    p = fma(p, r, -0x1.71547652b82fep-1); // -1/(2 ln 2)
    p = fma(p, r, 0x1.71547652b82fep+0);  //  1/ln 2
    return fma(r, p, PM_LOGC[i] + (double) e);
}

// 2^t as a double whose conversion to float is the wanted result. |t| >= 160 saturates to
// +-160 (2^160 converts to inf, 2^-160 to 0, like every larger magnitude), decided on the high
// word so that the test is a single integer comparison.
static inline double pm_exp2_core(double t)
{
    const uint64_t ut = pm_bits(t);
    if (((ut >> 32) & 0x7fffffffu) >= 0x40640000u)
        t = (ut >> 63) ? -160.0 : 160.0;
    const double shift = 6755399441055744.0; // 1.5 * 2^52: adding it rounds to nearest integer
    const double kd = (t * 64.0 + shift) - shift; // k = nearest integer to 64 t
    const double r = fma(kd, -0.015625, t);       // t - k/64, exact; |r| <= 1/128
    const int k = (int) kd;
    // 2^r = 1 + r (b1 + r (b2 + ... r b5)),  b_j = (ln 2)^j / j!
    double q = 0x1.5d87fe78a6731p-10;   // b5
    q = fma(q, r, 0x1.3b2ab6fba4e77p-7);  // b4
    q = fma(q, r, 0x1.c6b08d704a0cp-5);   // b3
    q = fma(q, r, 0x1.ebfbdff82c58fp-3);  // b2
    q = fma(q, r, 0x1.62e42fefa39efp-1);  // b1
    q = fma(q, r, 1.0);
    // 2^(k >> 6) * EXP2[k & 63]: add k >> 6 (floor division) to the table entry's exponent field
    const uint64_t tb = pm_bits(PM_EXP2[k & 63]) + ((uint64_t) (int64_t) (k >> 6) << 52);
    return q * pm_from_bits(tb);
}

static inline float pm_powf(float x, float y)
{
    if (y == 0.0f)
        return 1.0f;
    if (x != x || y != y)
        return NAN;
    if (x < 0.0f)
        return NAN; // GLSL: undefined; every caller guards with max(x, 0)
    if (x == 1.0f)
        return 1.0f;
    if (x == 0.0f)
        return y > 0.0f ? 0.0f : INFINITY;
    if (isinf(x))
        return y > 0.0f ? INFINITY : 0.0f;
    if (isinf(y))
        return ((x > 1.0f) == (y > 0.0f)) ? INFINITY : 0.0f;
    return (float) pm_exp2_core((double) y * pm_log2_core(x));
}

static inline float pm_expf(float x)
{
    if (x != x)
        return NAN;
    return (float) pm_exp2_core((double) x * 1.4426950408889634);
}

static inline float pm_logf(float x)
{
    if (x != x || x < 0.0f)
        return NAN;
    if (x == 0.0f)
        return -INFINITY;
    if (isinf(x))
        return INFINITY;
    return (float) (pm_log2_core(x) * 0.6931471805599453);
}

static inline float pm_log2f(float x)
{
    if (x != x || x < 0.0f)
        return NAN;
    if (x == 0.0f)
        return -INFINITY;
    if (isinf(x))
        return INFINITY;
    return (float) pm_log2_core(x);
}

// sine and cosine of a float of magnitude < 2^20
static inline void pm_sincosf(float x, float *sn, float *cs)
{
    const double xd = (double) x;
    const double shift = 6755399441055744.0;
    const double k = (xd * 0.6366197723675814 + shift) - shift;
    double r = fma(-k, 1.5707963267948966, xd);
    r = fma(-k, 6.123233995736766e-17, r);
    const double z = r * r;
    double ps = 2.8114572543455206e-15;
    ps = fma(ps, z, -7.647163731819816e-13);
    ps = fma(ps, z, 1.6059043836821613e-10);
    ps = fma(ps, z, -2.505210838544172e-08);
    ps = fma(ps, z, 2.7557319223985893e-06);
    ps = fma(ps, z, -0.0001984126984126984);
    ps = fma(ps, z, 0.008333333333333333);
    ps = fma(ps, z, -0.16666666666666666);
    const double sr = fma(r * z, ps, r);
    double pc = -1.5619206968586225e-16;
    pc = fma(pc, z, 4.779477332387385e-14);
    pc = fma(pc, z, -1.1470745597729725e-11);
    pc = fma(pc, z, 2.08767569878681e-09);
    pc = fma(pc, z, -2.755731922398589e-07);
    pc = fma(pc, z, 2.48015873015873e-05);
    pc = fma(pc, z, -0.001388888888888889);
    pc = fma(pc, z, 0.041666666666666664);
    pc = fma(pc, z, -0.5);
    const double cr = fma(z, pc, 1.0);
    const int q = (int) k & 3;
    const double s = (q & 1) ? cr : sr, c = (q & 1) ? sr : cr;
    *sn = (float) ((q & 2) ? -s : s);
    *cs = (float) (((q + 1) & 2) ? -c : c);
}

static inline float pm_atan2f(float y, float x)
{
    if (x != x || y != y)
        return NAN;
    const double a = fabs((double) y), b = fabs((double) x);
    double t;
    if (a == 0.0 && b == 0.0) {
        t = 0.0;
    } else if (isinf(a) || isinf(b)) {
        t = (isinf(a) && isinf(b)) ? 0.7853981633974483 : isinf(a) ? 1.5707963267948966 : 0.0;
    } else {
        const double lo = a < b ? a : b, hi = a < b ? b : a;
        double z, off;
        if (lo > 0.41421356237309503 * hi) {
            z = (lo - hi) / (lo + hi);
            off = 0.7853981633974483;
        } else {
            z = lo / hi;
            off = 0.0;
        }
        const double w = z * z;
        double p = -0.02857142857142857;
        p = fma(p, w, 0.030303030303030304);
        p = fma(p, w, -0.03225806451612903);
        p = fma(p, w, 0.034482758620689655);
        p = fma(p, w, -0.037037037037037035);
        p = fma(p, w, 0.04);
        p = fma(p, w, -0.043478260869565216);
        p = fma(p, w, 0.047619047619047616);
        p = fma(p, w, -0.05263157894736842);
        p = fma(p, w, 0.058823529411764705);
        p = fma(p, w, -0.06666666666666667);
        p = fma(p, w, 0.07692307692307693);
        p = fma(p, w, -0.09090909090909091);
        p = fma(p, w, 0.1111111111111111);
        p = fma(p, w, -0.14285714285714285);
        p = fma(p, w, 0.2);
        p = fma(p, w, -0.3333333333333333);
        t = off + fma(z * w, p, z);      // angle of (hi, lo) in [0, pi/4]
        if (a > b)
            t = 1.5707963267948966 - t;  // reflect about the diagonal
    }
    if (signbit(x))
        t = 3.141592653589793 - t;
    return (float) (signbit(y) ? -t : t);
}


// ---------------------------------------------------------------------------------------------
// Round 2: fp32 definitions used by the per-pixel hot path (deband's sin/cos, the hue angle of
// the gamut-mapping LUT). Only IEEE + - * / fma in a fixed order: same bits on any machine; the
// kernels restate them in csrc/plb_math.cuh (plm_atan2_fast, plm_sincos_fast). Accuracy against
// the real functions: atan2 <= 2e-7 absolute, sin/cos <= 1.2e-7 (tests/test_curves.py).
// orc_math_exact != 0 (orc_set_exact_math) selects the binary64 definitions above instead.
// ---------------------------------------------------------------------------------------------
extern int orc_math_exact;

// ratio folded to |z| <= tan(pi/8); atan z = z + z^3 P(z^2), P of degree 4 (minimax fit)
static inline float pf_atan2f(float y, float x)
{
    const float a = fabsf(y), b = fabsf(x);
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    if (orc_math_exact || !(hi > 0.0f && hi < INFINITY) || a != a || b != b)
        return pm_atan2f(y, x); // zeros, infinities, NaN
    const int fold = lo > 0.41421357f * hi;
    const float num = fold ? lo - hi : lo, den = fold ? lo + hi : hi;
    const float z = num / den, w = z * z;
    float p = -0.060332417f;
    p = fmaf(p, w, 0.10572814f);
    p = fmaf(p, w, -0.14239983f);
    p = fmaf(p, w, 0.19998211f);
    p = fmaf(p, w, -0.33333308f);
    float t = fmaf(z * w, p, z);
    if (fold)
        t += 0.78539819f;
    if (a > b)
        t = 1.5707964f - t;
    if (signbit(x))
        t = 3.1415927f - t;
    return signbit(y) ? -t : t;
}

// |x| <= 64: k = nearest quadrant, r = x - k pi/2 with a two-term constant
static inline void pf_sincosf(float x, float *sn, float *cs)
{
    if (orc_math_exact || !(fabsf(x) <= 64.0f)) {
        pm_sincosf(x, sn, cs);
        return;
    }
    const float k = floorf(fmaf(x, 0.63661975f, 0.5f));
    float r = fmaf(-k, 1.5707964f, x);
    r = fmaf(-k, -4.3711388e-8f, r);
    const float z = r * r;
    float ps = -0.00019503963f;
    ps = fmaf(ps, z, 0.0083320364f);
    ps = fmaf(ps, z, -0.16666651f);
    const float sr = fmaf(r * z, ps, r);
    float pc = 2.437988e-05f;
    pc = fmaf(pc, z, -0.0013886619f);
    pc = fmaf(pc, z, 0.041666616f);
    pc = fmaf(pc, z, -0.5f);
    const float cr = fmaf(z, pc, 1.0f);
    const int q = (int) k & 3;
    const float s = (q & 1) ? cr : sr, c = (q & 1) ? sr : cr;
    *sn = (q & 2) ? -s : s;
    *cs = ((q + 1) & 2) ? -c : c;
}

#endif
