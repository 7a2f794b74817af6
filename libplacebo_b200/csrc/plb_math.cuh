// Deterministic transcendental functions for the pixel kernels.
//
// GLSL leaves pow/exp/log/sin/cos/atan implementation-defined; libdevice's float versions are
// 1-4 ULP off the true value, and the PQ EOTF amplifies one ULP of pow() error ~500x. To make
// the colour chain reproducible bit for bit (against any IEEE-754 machine, and between runs on
// different GPU generations) these functions are DEFINED by short binary64 algorithms built
// only from +, *, /, fma and integer operations — all correctly rounded on sm_100a, which has
// full-rate-class FP64 pipes that otherwise sit idle in these kernels and issue beside the
// FP32/INT work. The float result is the correctly rounded value except when the exact result
// lies within ~1e-13 relative of a rounding boundary.
//
//   log2 x  : x = m 2^e, m in [sqrt(1/2), sqrt 2); s = (m-1)/(m+1); e + s P(s^2), 9 odd-series terms
//   exp2 t  : n = rint(t), r = t - n; 2^n Q(r), Taylor degree 12
//   pow     : exp2(y log2 x);  exp x = exp2(x log2 e);  log x = ln 2 * log2 x
//   sin/cos : quadrant reduction by pi/2 (two-term constant), Taylor degree 17/18
//   atan2   : ratio folded to |z| <= tan(pi/8), odd series with 18 terms, octant fix-ups
#pragma once
#include <cstdint>

#define PLM_DEV static __device__ __forceinline__

PLM_DEV double plm_log2_core(double x)
{
    // x: finite, > 0, converted from a float (hence a normal double)
    const long long u = __double_as_longlong(x);
    int e = (int) (u >> 52) - 1023;
    const long long mant = u & 0x000FFFFFFFFFFFFFll;
    const bool upper = mant > 0x6A09E667F3BCCll; // mantissa field of sqrt(2)
    e += upper ? 1 : 0;
    const double m = __longlong_as_double(mant | (upper ? 0x3FE0000000000000ll : 0x3FF0000000000000ll));
    const double s = __ddiv_rn(m - 1.0, m + 1.0);
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

PLM_DEV double plm_exp2_core(double t)
{
    t = t < -1100.0 ? -1100.0 : t > 1100.0 ? 1100.0 : t;
    const double shift = 6755399441055744.0; // 1.5 * 2^52
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
    const int ni = (int) n;
    const int n1 = ni / 2, n2 = ni - n1;
    const double s1 = __longlong_as_double((long long) (1023 + n1) << 52);
    const double s2 = __longlong_as_double((long long) (1023 + n2) << 52);
    return (q * s1) * s2;
}

PLM_DEV float plm_powf(float x, float y)
{
    // common case first: finite positive base, finite non-zero exponent
    if (x > 0.0f && x < __int_as_float(0x7f800000) && x != 1.0f && y != 0.0f &&
        fabsf(y) < __int_as_float(0x7f800000))
        return (float) plm_exp2_core((double) y * plm_log2_core((double) x));
    if (y == 0.0f)
        return 1.0f;
    if (x != x || y != y || x < 0.0f)
        return __int_as_float(0x7fc00000);
    if (x == 1.0f)
        return 1.0f;
    const float inf = __int_as_float(0x7f800000);
    if (x == 0.0f)
        return y > 0.0f ? 0.0f : inf;
    if (x == inf)
        return y > 0.0f ? inf : 0.0f;
    return ((x > 1.0f) == (y > 0.0f)) ? inf : 0.0f; // |y| = inf
}

PLM_DEV float plm_expf(float x)
{
    if (x != x)
        return x;
    return (float) plm_exp2_core((double) x * 1.4426950408889634);
}

PLM_DEV float plm_logf(float x)
{
    const float inf = __int_as_float(0x7f800000);
    if (x > 0.0f && x < inf)
        return (float) (plm_log2_core((double) x) * 0.6931471805599453);
    if (x == 0.0f)
        return -inf;
    if (x == inf)
        return inf;
    return __int_as_float(0x7fc00000);
}

// |x| < 2^20
PLM_DEV void plm_sincosf(float x, float *sn, float *cs)
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

PLM_DEV float plm_atan2f(float y, float x)
{
    if (x != x || y != y)
        return __int_as_float(0x7fc00000);
    const double a = fabs((double) y), b = fabs((double) x);
    const double dinf = __longlong_as_double(0x7FF0000000000000ll);
    double t;
    if (a == 0.0 && b == 0.0) {
        t = 0.0;
    } else if (a == dinf || b == dinf) {
        t = (a == dinf && b == dinf) ? 0.7853981633974483 : a == dinf ? 1.5707963267948966 : 0.0;
    } else {
        const double lo = a < b ? a : b, hi = a < b ? b : a;
        const bool fold = lo > 0.41421356237309503 * hi;
        const double z = fold ? __ddiv_rn(lo - hi, lo + hi) : __ddiv_rn(lo, hi);
        const double off = fold ? 0.7853981633974483 : 0.0;
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
        t = off + fma(z * w, p, z);
        if (a > b)
            t = 1.5707963267948966 - t;
    }
    if (__float_as_int(x) < 0)
        t = 3.141592653589793 - t;
    return (float) (__float_as_int(y) < 0 ? -t : t);
}
