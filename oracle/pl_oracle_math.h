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
//   log2(x):  x = m * 2^e with m in [sqrt(1/2), sqrt(2));  s = (m-1)/(m+1);
//             log2(x) = e + s * P9(s^2),   P9 = sum_{k<9} 2/((2k+1) ln 2) * s^(2k)
//   exp2(t):  t clamped to [-1100, 1100];  n = nearest integer;  r = t - n in [-1/2, 1/2];
//             exp2(t) = 2^n * Q12(r),      Q12 = sum_{k<=12} (ln 2)^k / k! * r^k
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

static inline double pm_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t pm_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

// log2 of a finite, strictly positive double that came from a float (so never subnormal as a
// double)
static inline double pm_log2_core(double x)
{
    uint64_t u = pm_bits(x);
    int e = (int) (u >> 52) - 1023;
    uint64_t mant = u & 0x000FFFFFFFFFFFFFull;
    // m in [1,2); move to [sqrt(1/2), sqrt(2)) — compare on the mantissa bits of sqrt(2)
    if (mant > 0x6A09E667F3BCCull) {
        e += 1;
        u = mant | 0x3FE0000000000000ull; // m/2 in [sqrt(1/2), 1)
    } else {
        u = mant | 0x3FF0000000000000ull;
    }
    const double m = pm_from_bits(u);
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 0.16972882833987804;
    p = fma(p, z, 0.19235933878519512);
    p = fma(p, z, 0.22195308321368667);
    p = fma(p, z, 0.2623081892525388);
    p = fma(p, z, 0.3205988979753252);
    p = fma(p, z, 0.4121985831111324);
    p = fma(p, z, 0.5770780163555853);
    p = fma(p, z, 0.9617966939259756);
    p = fma(p, z, 2.8853900817779268);
    return fma(s, p, (double) e);
}

// 2^t as a double (t finite)
static inline double pm_exp2_core(double t)
{
    t = t < -1100.0 ? -1100.0 : t > 1100.0 ? 1100.0 : t;
    const double shift = 6755399441055744.0; // 1.5 * 2^52: adding it rounds to nearest integer
    const double n = (t + shift) - shift;
    const double r = t - n;
    double q = 2.5678435993488206e-11;
    q = fma(q, r, 4.4455382718708116e-10);
    q = fma(q, r, 7.054911620801123e-09);
    q = fma(q, r, 1.01780860092397e-07);
    q = fma(q, r, 1.321548679014431e-06);
    q = fma(q, r, 1.5252733804059841e-05);
    q = fma(q, r, 0.0001540353039338161);
    q = fma(q, r, 0.0013333558146428443);
    q = fma(q, r, 0.009618129107628477);
    q = fma(q, r, 0.05550410866482158);
    q = fma(q, r, 0.24022650695910072);
    q = fma(q, r, 0.6931471805599453);
    q = fma(q, r, 1.0);
    // scale by 2^n in two exact steps so that |n| up to 1100 cannot overflow the exponent field
    const int ni = (int) n;
    const int n1 = ni / 2, n2 = ni - n1;
    const double s1 = pm_from_bits((uint64_t) (1023 + n1) << 52);
    const double s2 = pm_from_bits((uint64_t) (1023 + n2) << 52);
    return (q * s1) * s2;
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
    return (float) pm_exp2_core((double) y * pm_log2_core((double) x));
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
    return (float) (pm_log2_core((double) x) * 0.6931471805599453);
}

static inline float pm_log2f(float x)
{
    if (x != x || x < 0.0f)
        return NAN;
    if (x == 0.0f)
        return -INFINITY;
    if (isinf(x))
        return INFINITY;
    return (float) pm_log2_core((double) x);
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

#endif
