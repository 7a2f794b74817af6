// Host-side gamut-mapping 3DLUT generation in IPT/ICh space.
// Behavioural restatement of the reference's src/gamut_mapping.c: LUT driver (:366-443),
// gamut model and membership test (:246-364), boundary searches (:460-590), the mapping
// families (:711-990). Arithmetic keeps the reference's precision, operand order and search
// tolerances, because the LUT is quantised to 16 bits and compared bit-exactly
// (tests/test_t0_gamut_mapping.py). The 1024-entry PQ EOTF table the reference embeds as
// literals is regenerated here from its definition: entry i = EOTF((float) i / 1023) in
// double precision, rounded to 8 significant decimal digits (verified identical).
#define _GNU_SOURCE
#include "../internal.h"

#include <math.h>
#include <pthread.h>

#define MIX(a, b, x) ((x) * (b) + (1 - (x)) * (a))
static inline float fclampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

static inline float smoothstep(float e0, float e1, float x)
{
    if (e0 == e1)
        return x >= e0;
    x = (x - e0) / (e1 - e0);
    x = x < 0.0f ? 0.0f : x > 1.0f ? 1.0f : x;
    return x * x * (3.0f - 2.0f * x);
}

static const float PQ_M1 = 2610./4096 * 1./4, PQ_M2 = 2523./4096 * 128, PQ_C1 = 3424./4096,
                   PQ_C2 = 2413./4096 * 32, PQ_C3 = 2392./4096 * 32;

enum { PQ_TAB = 1024 };
static float pq_tab[PQ_TAB + 1];
static pthread_once_t pq_tab_once = PTHREAD_ONCE_INIT;

static void pq_tab_init(void)
{
    const double m1 = 2610./4096 / 4, m2 = 2523./4096 * 128, c1 = 3424./4096,
                 c2 = 2413./4096 * 32, c3 = 2392./4096 * 32;
    for (int i = 0; i < PQ_TAB; i++) {
        const double x = (float) i / 1023.0f;
        const double p = pow(x, 1 / m2);
        const double y = pow(fmax(p - c1, 0) / (c2 - c3 * p), 1 / m1);
        char buf[32];
        snprintf(buf, sizeof(buf), "%.7e", y);
        pq_tab[i] = strtof(buf, NULL);
    }
    pq_tab[PQ_TAB] = 1.0f;
}

// PQ code -> linear light normalised to 10000 nits, by table + linear interpolation
static inline float pq_eotf(float x)
{
    const float idxf = fminf(fmaxf(x, 0.0f), 1.0f) * (PQ_TAB - 1);
    const int ipart = floorf(idxf);
    const float fpart = idxf - ipart;
    return MIX(pq_tab[ipart], pq_tab[ipart + 1], fpart);
}

static inline float pq_oetf(float x)
{
    x = powf(fmaxf(x, 0.0f), PQ_M1);
    x = (PQ_C1 + PQ_C2 * x) / (1.0f + PQ_C3 * x);
    return powf(x, PQ_M2);
}

struct rgb { float R, G, B; };
struct ipt { float I, P, T; };
struct ich { float I, C, h; };

static inline struct ich to_ich(struct ipt c)
{
    return (struct ich) { c.I, sqrtf(c.P * c.P + c.T * c.T), atan2f(c.T, c.P) };
}

static inline struct ipt to_ipt(struct ich c)
{
    return (struct ipt) { c.I, c.C * cosf(c.h), c.C * sinf(c.h) };
}

// An RGB gamut seen from IPT: conversion matrices, luminance range (PQ) and the matching
// RGB range (linear, with a little slack), plus a one-entry memo of the last cusp search
struct gamut {
    pl_matrix3x3 lms2rgb, rgb2lms;
    float min_luma, max_luma;
    float min_rgb, max_rgb;
    struct ich *cusp_memo;
};

struct memo { struct ich src, dst; };

static void setup_gamuts(struct gamut *dst, struct gamut *src, struct memo *memo,
                         const struct pl_gamut_map_params *p)
{
    const float eps = 1e-6;
    memset(memo, 0, sizeof(*memo));
    const struct gamut base = {
        .min_luma = p->min_luma,
        .max_luma = p->max_luma,
        .min_rgb = pq_eotf(p->min_luma) - eps,
        .max_rgb = pq_eotf(p->max_luma) + eps,
    };
    if (dst) {
        *dst = base;
        dst->lms2rgb = dst->rgb2lms = pl_ipt_rgb2lms(&p->output_gamut);
        dst->cusp_memo = &memo->dst;
        pl_matrix3x3_invert(&dst->lms2rgb);
    }
    if (src) {
        *src = base;
        src->lms2rgb = src->rgb2lms = pl_ipt_rgb2lms(&p->input_gamut);
        src->cusp_memo = &memo->src;
        pl_matrix3x3_invert(&src->lms2rgb);
    }
}

static inline struct ipt rgb_to_ipt(struct rgb c, const struct gamut *g)
{
    const float (*m)[3] = g->rgb2lms.m;
    const float L = m[0][0] * c.R + m[0][1] * c.G + m[0][2] * c.B;
    const float M = m[1][0] * c.R + m[1][1] * c.G + m[1][2] * c.B;
    const float S = m[2][0] * c.R + m[2][1] * c.G + m[2][2] * c.B;
    const float Lp = pq_oetf(L), Mp = pq_oetf(M), Sp = pq_oetf(S);
    return (struct ipt) {
        .I = 0.4000f * Lp + 0.4000f * Mp + 0.2000f * Sp,
        .P = 4.4550f * Lp - 4.8510f * Mp + 0.3960f * Sp,
        .T = 0.8056f * Lp + 0.3572f * Mp - 1.1628f * Sp,
    };
}

static inline void ipt_to_lmspq(struct ipt c, float *Lp, float *Mp, float *Sp)
{
    *Lp = c.I + 0.0975689f * c.P + 0.205226f * c.T;
    *Mp = c.I - 0.1138760f * c.P + 0.133217f * c.T;
    *Sp = c.I + 0.0326151f * c.P - 0.676887f * c.T;
}

static inline struct rgb lms_to_rgb(float L, float M, float S, const struct gamut *g)
{
    const float (*m)[3] = g->lms2rgb.m;
    return (struct rgb) {
        .R = m[0][0] * L + m[0][1] * M + m[0][2] * S,
        .G = m[1][0] * L + m[1][1] * M + m[1][2] * S,
        .B = m[2][0] * L + m[2][1] * M + m[2][2] * S,
    };
}

static inline struct rgb ipt_to_rgb(struct ipt c, const struct gamut *g)
{
    float Lp, Mp, Sp;
    ipt_to_lmspq(c, &Lp, &Mp, &Sp);
    return lms_to_rgb(pq_eotf(Lp), pq_eotf(Mp), pq_eotf(Sp), g);
}

static inline bool inside(struct ipt c, const struct gamut *g)
{
    float Lp, Mp, Sp;
    ipt_to_lmspq(c, &Lp, &Mp, &Sp);
    if (Lp < g->min_luma || Lp > g->max_luma || Mp < g->min_luma || Mp > g->max_luma ||
        Sp < g->min_luma || Sp > g->max_luma)
        return false;
    const struct rgb v = lms_to_rgb(pq_eotf(Lp), pq_eotf(Mp), pq_eotf(Sp), g);
    return v.R >= g->min_rgb && v.R <= g->max_rgb && v.G >= g->min_rgb && v.G <= g->max_rgb &&
           v.B >= g->min_rgb && v.B <= g->max_rgb;
}

static const float max_delta = 5e-5f;

// Largest in-gamut chroma at (I, h), searched by bisection inside [Cmin, Cmax]
static inline struct ich boundary_chroma(float I, float h, float Cmin, float Cmax,
                                         const struct gamut *g)
{
    if (I <= g->min_luma)
        return (struct ich) { g->min_luma, 0, h };
    if (I >= g->max_luma)
        return (struct ich) { g->max_luma, 0, h };
    const float tol = I * max_delta;
    struct ich res = { I, (Cmin + Cmax) / 2, h };
    do {
        if (inside(to_ipt(res), g))
            Cmin = res.C;
        else
            Cmax = res.C;
        res.C = (Cmin + Cmax) / 2;
    } while (Cmax - Cmin > tol);
    return res;
}

// The cusp (most saturated point) of the gamut at a hue: golden-section search over I
static inline struct ich cusp(float hue, const struct gamut *g)
{
    if (g->cusp_memo->I && fabsf(g->cusp_memo->h - hue) < 1e-3)
        return *g->cusp_memo;
    static const float invphi = 0.6180339887498948f, invphi2 = 0.38196601125010515f;
    struct ich lo = { .I = g->min_luma, .h = hue }, hi = { .I = g->max_luma, .h = hue };
    float de = hi.I - lo.I;
    struct ich a = { .I = lo.I + invphi2 * de }, b = { .I = lo.I + invphi * de };
    a = boundary_chroma(a.I, hue, 0.0f, 0.5f, g);
    b = boundary_chroma(b.I, hue, 0.0f, 0.5f, g);
    while (de > max_delta) {
        de *= invphi;
        if (a.C > b.C) {
            hi = b;
            b = a;
            a.I = lo.I + invphi2 * de;
            a = boundary_chroma(a.I, hue, lo.C - max_delta, 0.5f, g);
        } else {
            lo = a;
            a = b;
            b.I = lo.I + invphi * de;
            b = boundary_chroma(b.I, hue, hi.C - max_delta, 0.5f, g);
        }
    }
    const struct ich peak = a.C > b.C ? a : b;
    *g->cusp_memo = peak;
    return peak;
}

// move x of the way from the cusp luminance towards the colour, with chroma scaled by x
static inline struct ich toward_cusp(struct ich c, float x, float gamma, float base)
{
    return (struct ich) { base + (c.I - base) * powf(x, gamma), c.C * x, c.h };
}

// Colorimetric clip: pull out-of-gamut colours towards the cusp luminance along a curve whose
// exponent grows with brightness and saturation; gamma == 0 is a pure chroma reduction
static inline struct ipt clip_to_gamut(struct ipt c, float gamma, const struct gamut *g)
{
    if (c.I <= g->min_luma)
        return (struct ipt) { .I = g->min_luma };
    if (inside(c, g))
        return c;
    struct ich ich = to_ich(c);
    if (!gamma)
        return to_ipt(boundary_chroma(ich.I, ich.h, 0.0f, ich.C, g));
    const float tol = fmaxf(ich.I * max_delta, 1e-7f);
    const struct ich peak = cusp(ich.h, g);
    const float Irel = fmaxf((ich.I - g->min_luma) / (peak.I - g->min_luma), 0.0f);
    gamma = gamma * powf(Irel, 3) * fminf(ich.C / peak.C, 1.0f);
    float lo = 0.0f, hi = 1.0f, x = 0.5f;
    do {
        if (inside(to_ipt(toward_cusp(ich, x, gamma, peak.I)), g))
            lo = x;
        else
            hi = x;
        x = (lo + hi) / 2.0f;
    } while (hi - lo > tol);
    return to_ipt(toward_cusp(ich, x, gamma, peak.I));
}

// Mobius-style knee compressing [0, source] into [0, target]
static float soft_knee(float value, float source, float target,
                       const struct pl_gamut_map_constants *c)
{
    if (!target)
        return 0.0f;
    const float peak = source / target;
    const float x = fminf(value / target, peak);
    const float j = c->softclip_knee;
    if (x <= j || peak <= 1.0)
        return value;
    const float a = -j*j * (peak - 1.0f) / (j*j - 2.0f * j + peak);
    const float b = (j*j - 2.0f * j * peak + peak) / fmaxf(1e-6f, peak - 1.0f);
    const float scale = (b*b + 2.0f * b*j + j*j) / (b - a);
    return scale * (x + a) / (x + b) * target;
}

// ---- LUT iteration -------------------------------------------------------------------------------

#define LUT_FLOATS(p) ((p)->lut_size_I * (p)->lut_size_C * (p)->lut_size_h * (p)->lut_stride)
#define EACH(lut, c) \
    for (struct ipt *it_ = (struct ipt *) (lut), *end_ = (struct ipt *) ((lut) + LUT_FLOATS(p)), c; \
         it_ < end_ && (c = *it_, 1); \
         *it_ = c, it_ = (struct ipt *) ((float *) it_ + p->lut_stride))

static void map_none(float *lut, const struct pl_gamut_map_params *p) { (void) lut; (void) p; }

static void map_perceptual(float *lut, const struct pl_gamut_map_params *p)
{
    const struct pl_gamut_map_constants *k = &p->constants;
    struct memo memo;
    struct gamut dst, src;
    setup_gamuts(&dst, &src, &memo, p);
    EACH(lut, c) {
        const struct ich ich = to_ich(c);
        const struct ich src_peak = cusp(ich.h, &src), dst_peak = cusp(ich.h, &dst);
        // blend towards the "same RGB, other primaries" colour as saturation rises
        const struct ipt mapped = rgb_to_ipt(ipt_to_rgb(c, &src), &dst);
        const float maxC = fmaxf(src_peak.C, dst_peak.C);
        float w = smoothstep(k->perceptual_deadzone, 1.0f, ich.C / maxC);
        w *= k->perceptual_strength;
        c.I = MIX(c.I, mapped.I, w);
        c.P = MIX(c.P, mapped.P, w);
        c.T = MIX(c.T, mapped.T, w);
        struct rgb v = ipt_to_rgb(c, &dst);
        const float top = fmaxf(v.R, fmaxf(v.G, v.B));
        v.R = fmaxf(soft_knee(v.R, top, dst.max_rgb, k), dst.min_rgb);
        v.G = fmaxf(soft_knee(v.G, top, dst.max_rgb, k), dst.min_rgb);
        v.B = fmaxf(soft_knee(v.B, top, dst.max_rgb, k), dst.min_rgb);
        c = rgb_to_ipt(v, &dst);
    }
}

// --- hue shift spline used by the soft-clip mapping ---

enum { HS = 12, HN = HS + 2 };

struct hue_node { float hue, delta; };

struct hue_spline {
    float dh[HN], slope[HN], K[HN];
    float last_hue, last_shift;
    struct hue_node node[HN];
};

static int node_cmp(const void *a, const void *b)
{
    const float fa = *(const float *) a, fb = *(const float *) b;
    return (fa > fb) - (fa < fb);
}

static float wrap_pi(float h)
{
    if (h > M_PI)
        return h - 2 * M_PI;
    if (h < -M_PI)
        return h + 2 * M_PI;
    return h;
}

// natural cubic spline through the hue displacement of 12 reference colours
static void hue_spline_init(struct hue_spline *s, const struct gamut *src, const struct gamut *dst)
{
    const float O = pq_eotf(src->min_luma), X = pq_eotf(src->max_luma);
    const float M = (O + X) / 2.0f;
    const struct rgb ref[HS] = {
        {X, O, O}, {O, X, O}, {O, O, X}, {O, X, X}, {X, O, X}, {X, X, O},
        {O, X, M}, {X, O, M}, {X, M, O}, {O, M, X}, {M, O, X}, {M, X, O},
    };
    memset(s, 0, sizeof(*s));
    for (int i = 0; i < HS; i++) {
        const struct ich a = to_ich(rgb_to_ipt(ref[i], src)), b = to_ich(rgb_to_ipt(ref[i], dst));
        s->node[i + 1].hue = a.h;
        s->node[i + 1].delta = wrap_pi(b.h - a.h);
    }
    qsort(s->node + 1, HS, sizeof(*s->node), node_cmp);
    // periodic extension by one node on each side
    s->node[0] = s->node[HS];
    s->node[HS + 1] = s->node[1];
    s->node[0].hue -= 2 * M_PI;
    s->node[HS + 1].hue += 2 * M_PI;

    float tri[HN][HN] = {{0}};
    for (int i = HN - 1; i > 0; i--) {
        s->dh[i - 1] = s->node[i].hue - s->node[i - 1].hue;
        s->slope[i] = (s->node[i].delta - s->node[i - 1].delta) / s->dh[i - 1];
    }
    for (int i = 1; i < HN - 1; i++) {
        tri[i][i] = 2 * (s->dh[i - 1] + s->dh[i]);
        if (i != 1)
            tri[i][i - 1] = tri[i - 1][i] = s->dh[i - 1];
        tri[i][HN - 1] = 6 * (s->slope[i + 1] - s->slope[i]);
    }
    for (int i = 1; i < HN - 2; i++) {
        const float q = (tri[i + 1][i] / tri[i][i]);
        for (int j = 1; j <= HN - 1; j++)
            tri[i + 1][j] -= q * tri[i][j];
    }
    for (int i = HN - 2; i > 0; i--) {
        float sum = 0.0f;
        for (int j = i; j <= HN - 2; j++)
            sum += tri[i][j] * s->K[j];
        s->K[i] = (tri[i][HN - 1] - sum) / tri[i][i];
    }
    s->last_hue = -10.0f;
}

static struct ich hue_spline_apply(struct hue_spline *s, struct ich c)
{
    if (!(fabsf(c.h - s->last_hue) < 1e-6f)) {
        for (int i = 0; i < HN - 1; i++) {
            if (s->node[i + 1].hue > c.h) {
                const float a = (s->K[i + 1] - s->K[i]) / (6 * s->dh[i]);
                const float b = s->K[i] / 2;
                const float cc = s->slope[i + 1] -
                                 (2 * s->dh[i] * s->K[i] + s->K[i + 1] * s->dh[i]) / 6;
                const float d = s->node[i].delta;
                const float x = c.h - s->node[i].hue;
                const float delta = ((a * x + b) * x + cc) * x + d;
                s->last_shift = c.h + delta;
                s->last_hue = c.h;
                break;
            }
        }
    }
    return (struct ich) { c.I, c.C, s->last_shift };
}

static void map_softclip(float *lut, const struct pl_gamut_map_params *p)
{
    const struct pl_gamut_map_constants *k = &p->constants;
    struct memo memo_pre, memo_post;
    struct gamut dst_pre, src_pre, src_post, dst_post;
    struct hue_spline hs;
    setup_gamuts(&dst_pre, &src_pre, &memo_pre, p);
    setup_gamuts(&dst_post, &src_post, &memo_post, p);
    hue_spline_init(&hs, &src_pre, &dst_pre);
    EACH(lut, c) {
        const struct gamut *src = &src_pre, *dst = &dst_pre;
        if (c.I <= dst->min_luma) {
            c.P = c.T = 0.0f;
            continue;
        }
        struct ich ich = to_ich(c);
        if (ich.C <= 1e-2f)
            continue;
        float margin = 1.0f;
        const struct ich shifted = hue_spline_apply(&hs, ich);
        if (fabsf(shifted.h - ich.h) >= 1e-3f) {
            const struct ich src_border = boundary_chroma(ich.I, ich.h, 0.0f, 0.5f, src);
            const struct ich dst_border = boundary_chroma(ich.I, ich.h, 0.0f, 0.5f, dst);
            const float w = smoothstep(dst_border.C * k->softclip_knee, src_border.C, ich.C);
            ich.h = MIX(ich.h, shifted.h, w);
            src = &src_post;
            dst = &dst_post;
            const struct ich shift_border = boundary_chroma(ich.I, ich.h, 0.0f, 0.5f, src);
            margin *= fmaxf(1.0f, src_border.C / shift_border.C);
        }
        const struct ich source = cusp(ich.h, src), target = cusp(ich.h, dst);
        const struct ich border = boundary_chroma(ich.I, ich.h, 0.0f, target.C, dst);
        const float chroma = MIX(target.C, border.C, k->softclip_desat);
        ich.C = soft_knee(ich.C, margin * source.C, chroma, k);
        const struct ich saturated = { ich.I, chroma, ich.h };
        const struct rgb peak = ipt_to_rgb(to_ipt(saturated), dst);
        struct rgb v = ipt_to_rgb(to_ipt(ich), dst);
        v.R = fmaxf(soft_knee(v.R, peak.R, dst->max_rgb, k), dst->min_rgb);
        v.G = fmaxf(soft_knee(v.G, peak.G, dst->max_rgb, k), dst->min_rgb);
        v.B = fmaxf(soft_knee(v.B, peak.B, dst->max_rgb, k), dst->min_rgb);
        c = rgb_to_ipt(v, dst);
    }
}

static void map_relative(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst;
    setup_gamuts(&dst, NULL, &memo, p);
    EACH(lut, c)
        c = clip_to_gamut(c, p->constants.colorimetric_gamma, &dst);
}

static void map_desaturate(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst;
    setup_gamuts(&dst, NULL, &memo, p);
    EACH(lut, c)
        c = clip_to_gamut(c, 0.0f, &dst);
}

static void map_saturation(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst, src;
    setup_gamuts(&dst, &src, &memo, p);
    EACH(lut, c)
        c = rgb_to_ipt(ipt_to_rgb(c, &src), &dst);
}

static void map_absolute(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst;
    setup_gamuts(&dst, NULL, &memo, p);
    const pl_matrix3x3 m = pl_get_adaptation_matrix(p->output_gamut.white, p->input_gamut.white);
    EACH(lut, c) {
        struct rgb v = ipt_to_rgb(c, &dst);
        pl_matrix3x3_apply(&m, (float *) &v);
        c = rgb_to_ipt(v, &dst);
        c = clip_to_gamut(c, p->constants.colorimetric_gamma, &dst);
    }
}

static void map_highlight(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst;
    setup_gamuts(&dst, NULL, &memo, p);
    EACH(lut, c) {
        if (!inside(c, &dst)) {
            c.I = fminf(c.I + 0.1f, 1.0f);
            c.P = fclampf(-1.2f * c.P, -0.5f, 0.5f);
            c.T = fclampf(-1.2f * c.T, -0.5f, 0.5f);
        }
    }
}

static void map_linear(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst, src;
    setup_gamuts(&dst, &src, &memo, p);
    float gain = 1.0f;
    for (float hue = -M_PI; hue < M_PI; hue += 0.1f)
        gain = fminf(gain, cusp(hue, &dst).C / cusp(hue, &src).C);
    EACH(lut, c) {
        struct ich ich = to_ich(c);
        ich.C *= gain;
        c = to_ipt(ich);
    }
}

static void map_darken(float *lut, const struct pl_gamut_map_params *p)
{
    struct memo memo;
    struct gamut dst, src;
    setup_gamuts(&dst, &src, &memo, p);
    static const struct rgb corners[6] = {
        {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0},
    };
    float gain = 1.0f;
    for (int i = 0; i < 6; i++) {
        const struct rgb v = ipt_to_rgb(rgb_to_ipt(corners[i], &src), &dst);
        const float top = PLB_MAX(PLB_MAX(v.R, v.G), v.B);
        gain = fminf(gain, 1.0 / top);
    }
    EACH(lut, c) {
        struct rgb v = ipt_to_rgb(c, &dst);
        v.R *= gain;
        v.G *= gain;
        v.B *= gain;
        c = rgb_to_ipt(v, &dst);
        c = clip_to_gamut(c, p->constants.colorimetric_gamma, &dst);
    }
}

#define GMF(sym, nm, desc, fn, bidi) \
    const struct pl_gamut_map_function sym = { .name = nm, .description = desc, .map = fn, \
                                               .bidirectional = bidi }

GMF(pl_gamut_map_clip, "clip", "No gamut mapping (hard clip)", map_none, false);
GMF(pl_gamut_map_perceptual, "perceptual", "Perceptual mapping", map_perceptual, true);
GMF(pl_gamut_map_softclip, "softclip", "Soft clipping", map_softclip, false);
GMF(pl_gamut_map_relative, "relative", "Colorimetric clip", map_relative, false);
GMF(pl_gamut_map_saturation, "saturation", "Saturation mapping", map_saturation, true);
GMF(pl_gamut_map_absolute, "absolute", "Absolute colorimetric clip", map_absolute, false);
GMF(pl_gamut_map_desaturate, "desaturate", "Desaturating clip", map_desaturate, false);
GMF(pl_gamut_map_darken, "darken", "Darken and clip", map_darken, false);
GMF(pl_gamut_map_highlight, "highlight", "Highlight out-of-gamut pixels", map_highlight, false);
GMF(pl_gamut_map_linear, "linear", "Linear desaturate", map_linear, false);

const struct pl_gamut_map_function * const pl_gamut_map_functions[] = {
    &pl_gamut_map_clip, &pl_gamut_map_perceptual, &pl_gamut_map_softclip, &pl_gamut_map_relative,
    &pl_gamut_map_saturation, &pl_gamut_map_absolute, &pl_gamut_map_desaturate,
    &pl_gamut_map_darken, &pl_gamut_map_highlight, &pl_gamut_map_linear, NULL
};
const int pl_num_gamut_map_functions = PLB_ARRAY_SIZE(pl_gamut_map_functions) - 1;

const struct pl_gamut_map_function *pl_find_gamut_map_function(const char *name)
{
    for (int i = 0; name && i < pl_num_gamut_map_functions; i++) {
        if (strcmp(name, pl_gamut_map_functions[i]->name) == 0)
            return pl_gamut_map_functions[i];
    }
    return NULL;
}

// ---- public driver ---------------------------------------------------------------------------------

static const struct pl_gamut_map_function *fn_of(const struct pl_gamut_map_params *p)
{
    return p->function ? p->function : &pl_gamut_map_clip;
}

static void sanitize(struct pl_gamut_map_constants *c)
{
    c->perceptual_deadzone = fclampf(c->perceptual_deadzone, 0.0f, 1.0f);
    c->perceptual_strength = fclampf(c->perceptual_strength, 0.0f, 1.0f);
    c->colorimetric_gamma = fclampf(c->colorimetric_gamma, 0.0f, 10.0f);
    c->softclip_knee = fclampf(c->softclip_knee, 0.0f, 1.0f);
    c->softclip_desat = fclampf(c->softclip_desat, 0.0f, 1.0f);
}

bool pl_gamut_map_params_equal(const struct pl_gamut_map_params *a, const struct pl_gamut_map_params *b)
{
    return a->function == b->function && a->min_luma == b->min_luma &&
           a->max_luma == b->max_luma && a->lut_size_I == b->lut_size_I &&
           a->lut_size_C == b->lut_size_C && a->lut_size_h == b->lut_size_h &&
           a->lut_stride == b->lut_stride &&
           !memcmp(&a->constants, &b->constants, sizeof(a->constants)) &&
           pl_raw_primaries_equal(&a->input_gamut, &b->input_gamut) &&
           pl_raw_primaries_equal(&a->output_gamut, &b->output_gamut);
}

bool pl_gamut_map_params_noop(const struct pl_gamut_map_params *p)
{
    if (fn_of(p)->map == map_none)
        return true;
    const struct pl_raw_primaries src = p->input_gamut, dst = p->output_gamut;
    if (!pl_primaries_compatible(&dst, &src))
        return true;
    bool need = !pl_primaries_superset(&dst, &src);
    need |= !(src.white.x == dst.white.x && src.white.y == dst.white.y);
    if (fn_of(p)->bidirectional)
        need |= !pl_raw_primaries_equal(&dst, &src);
    return !need;
}

struct slab {
    const struct pl_gamut_map_params *p;
    float *out;
    int h0, count;
};

// Fill one slab of hue planes with the identity ICh grid, then map it in place
static void *slab_run(void *priv)
{
    const struct slab *s = priv;
    const struct pl_gamut_map_params *p = s->p;
    float *o = s->out;
    for (int h = s->h0; h < s->h0 + s->count; h++) {
        for (int C = 0; C < p->lut_size_C; C++) {
            for (int I = 0; I < p->lut_size_I; I++) {
                const float Ix = (float) I / (p->lut_size_I - 1);
                const float Cx = (float) C / (p->lut_size_C - 1);
                const float hx = (float) h / (p->lut_size_h - 1);
                const struct ipt c = to_ipt((struct ich) {
                    .I = MIX(p->min_luma, p->max_luma, Ix),
                    .C = MIX(0.0f, 0.5f, Cx),
                    .h = MIX(-M_PI, M_PI, hx),
                });
                o[0] = c.I;
                o[1] = c.P;
                o[2] = c.T;
                o += p->lut_stride;
            }
        }
    }
    struct pl_gamut_map_params q = *p;
    sanitize(&q.constants);
    q.lut_size_h = s->count;
    fn_of(p)->map(s->out, &q);
    return NULL;
}

void pl_gamut_map_generate(float *out, const struct pl_gamut_map_params *p)
{
    pthread_once(&pq_tab_once, pq_tab_init);
    enum { MAX_SLABS = 32 };
    struct slab slabs[MAX_SLABS];
    pthread_t th[MAX_SLABS];
    bool started[MAX_SLABS] = {0};
    const int per = (p->lut_size_h + MAX_SLABS - 1) / MAX_SLABS;
    const int n = (p->lut_size_h + per - 1) / per;
    for (int i = 0; i < n; i++) {
        const int h0 = i * per;
        slabs[i] = (struct slab) { p, out, h0, PLB_MIN(per, p->lut_size_h - h0) };
        out += slabs[i].count * p->lut_size_C * p->lut_size_I * p->lut_stride;
    }
    for (int i = 0; i < n; i++)
        started[i] = pthread_create(&th[i], NULL, slab_run, &slabs[i]) == 0;
    for (int i = 0; i < n; i++) {
        if (started[i])
            pthread_join(th[i], NULL);
        else
            slab_run(&slabs[i]);
    }
}

void pl_gamut_map_sample(float x[3], const struct pl_gamut_map_params *p)
{
    pthread_once(&pq_tab_once, pq_tab_init);
    struct pl_gamut_map_params q = *p;
    sanitize(&q.constants);
    q.lut_size_I = q.lut_size_C = q.lut_size_h = 1;
    q.lut_stride = 3;
    fn_of(p)->map(x, &q);
}
