// Transfer curves: the composite transfer functions of the colour chain (PQ EOTF / OETF, the
// pure powers of BT.1886 and the gamma curves) as piecewise cubics indexed by the bits of the
// float argument. Shared by the host generator (curves.c) and the kernels (plb_math.cuh).
//
// Why: the reference's GLSL evaluates e.g. the PQ EOTF (colorspace.c:662-668) as
// pow -> rational -> pow with implementation-defined pow(). Round 1 defined pow() as a correctly
// rounded binary64 algorithm (55 instructions, 21 powers per HDR pixel: 1.8 % of the HBM
// roofline). The tolerance of the path is +-1 LSB of the 16-bit output, so the composite
// function f itself is now DEFINED as
//
//     f(x) ~ c0 + u (c1 + u (c2 + u c3)),  u in [-1/2, 1/2) the position inside the segment,
//
// one segment per 2^-M of a binade (segment index = float bits >> (23 - M)), with coefficients
// that interpolate the real-valued function (evaluated in binary64 from the same constants the
// GLSL text carries) at the four Chebyshev nodes of the segment. Evaluation costs ~12
// instructions (three FFMA, one 16-byte table fetch, integer bit slicing) and is within ~1 ULP of
// the real function — closer to it than the pow/divide/pow chain in fp32 could ever be, because
// the cancellation in (x^(1/m2) - c1) never happens in fp32. Arguments outside the tabulated
// domain (negative, NaN, beyond `hi`) take the round-1 exact path, so the case analysis of
// pow() is unchanged. PLB_EXACT_MATH=1 empties every domain (regression switch).
#ifndef PLB_CURVES_H_
#define PLB_CURVES_H_

enum plb_curve_id {
    PLB_CV_PQ_EOTF_SDR = 0, // from_pq * PQ EOTF(x), GLSL literal constants (linearize, colour map)
    PLB_CV_PQ_EOTF,         // PQ EOTF(x)                       (ICtCp / Dolby Vision decode)
    PLB_CV_PQ_OETF_SDR,     // PQ OETF(to_pq * x)               (colour map)
    PLB_CV_PQ_OETF_DELIN,   // PQ OETF(x / from_pq)             (delinearize)
    PLB_CV_PQ_OETF,         // PQ OETF(x)                       (ICtCp encode, feature map)
    PLB_CV_PQX_OETF,        // PQ OETF(x), exact float constants (peak detection, :1279-1288)
    PLB_CV_POW_24,          // x^2.4        BT.1886 / gamma 2.4 linearize
    PLB_CV_POW_INV24,       // x^(1/2.4)    BT.1886 / gamma 2.4 delinearize
    PLB_CV_POW_22,          // x^2.2
    PLB_CV_POW_INV22,       // x^(1/2.2)
    PLB_CV_COUNT,
};

// segments per binade = 2^M, fixed per curve so that the kernels shift by immediates
#define PLB_CV_MBITS(id) \
    ((id) <= PLB_CV_PQ_EOTF ? 7 : (id) <= PLB_CV_PQX_OETF ? 4 : \
     ((id) == PLB_CV_POW_24 || (id) == PLB_CV_POW_22) ? 6 : 4)

// Domain of each curve: [2^EMIN, float with bits HIBITS). Compile-time so that the specialised
// kernels slice the argument with immediates (plm_curve_s); the generator uses the same numbers.
//   PQ EOTF: zero below c1^m2 = 2^-20.4; the denominator c2 - c3 x^(1/m2) has its root at 1.99
#define PLB_CV_EMIN(id) \
    ((id) <= PLB_CV_PQ_EOTF ? -21 : (id) <= PLB_CV_PQ_OETF_DELIN ? -58 : (id) <= PLB_CV_PQX_OETF ? -64 : \
     ((id) == PLB_CV_POW_24 || (id) == PLB_CV_POW_22) ? -20 : -96)
#define PLB_CV_HIBITS(id) \
    ((id) <= PLB_CV_PQ_EOTF ? 0x3FC00000u /* 1.5 */ : (id) <= PLB_CV_PQ_OETF_DELIN ? 0x43800000u /* 256 */ : \
     (id) == PLB_CV_PQ_OETF ? 0x41000000u /* 8 */ : (id) == PLB_CV_PQX_OETF ? 0x40000000u /* 2 */ : \
     ((id) == PLB_CV_POW_24 || (id) == PLB_CV_POW_22) ? 0x41000000u : 0x41800000u /* 16 */)
// first segment index - 1 (= index of the constant segment below the domain) and segment count
#define PLB_CV_BASE(id) (((PLB_CV_EMIN(id) + 127) << PLB_CV_MBITS(id)) - 1)
#define PLB_CV_NSEG(id) ((int) (PLB_CV_HIBITS(id) >> (23 - PLB_CV_MBITS(id))) - (PLB_CV_BASE(id) + 1))
// table entries: [0] constant below the domain, [1 .. nseg] the segments, [nseg + 1] unused
// padding that clamped out-of-domain indices land on
#define PLB_CV_ENTRIES(id) (PLB_CV_NSEG(id) + 2)

// Device-side descriptor. Segment 0 is the constant f(0) used for every argument below the
// domain; segment i >= 1 covers float bits [(base + i) << (23 - M), (base + i + 1) << (23 - M)).
struct plb_curve {
    const void *coef;   // float4[nseg + 2]: (c0, c1, c2, c3)
    int base;
    int nseg;           // -1: empty domain (everything takes the exact path)
};

// All tables live in ONE device allocation, curve `id` starting at entry PLB_CV_OFFSET(id), so
// that a specialised kernel needs a single base pointer and immediate offsets.
static inline
#ifdef __CUDACC__
__host__ __device__ constexpr
#endif
int plb_cv_offset(int id)
{
    int off = 0;
    for (int j = 0; j < id; j++)
        off += PLB_CV_ENTRIES(j);
    return off;
}

// Host-side table as generated
struct plb_curve_host {
    float *coef;        // 4 * (nseg + 1) floats, malloc'ed
    int base, nseg, mbits;
};

#endif
