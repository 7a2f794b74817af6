/*
 * TEST INFRASTRUCTURE — see pl_oracle.h. Part 2: colour decoding/encoding, transfer
 * functions, peak detection, colour mapping, dithering, error diffusion.
 *
 * Constants: where the reference prints a constant into the GLSL text with "%f"
 * (6 decimals) the shader really computes with the truncated decimal, so the oracle does
 * too (glsl6()); where it passes an exact float (sh_const / SH_FLOAT / ${const float:..})
 * the oracle uses the exact float. "1.0/%f" is folded in double and rounded once, like
 * glslang's constant folder.
 */
#define _GNU_SOURCE
#include "pl_oracle.h"

#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include "pl_oracle_math.h"
#include "pl_oracle_curve.h"
#include <string.h>

static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float smoothstepf(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// value of the decimal text that printf("%f") produces
static double glsl6(double x)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "%f", x);
    return strtod(buf, NULL);
}

// src/colorspace.h:23-33
static const float PQ_M1 = 2610./4096 * 1./4,
                   PQ_M2 = 2523./4096 * 128,
                   PQ_C1 = 3424./4096,
                   PQ_C2 = 2413./4096 * 32,
                   PQ_C3 = 2392./4096 * 32;
static const float HLG_A = 0.17883277, HLG_B = 0.28466892, HLG_C = 0.55991073;
static const float VLOG_B = 0.00873, VLOG_C = 0.241514, VLOG_D = 0.598206;
static const float SLOG_A = 0.432699, SLOG_B = 0.037584, SLOG_C = 0.616596 + 0.03,
                   SLOG_P = 3.538813, SLOG_Q = 0.030001, SLOG_K2 = 155.0 / 219.0;
#define SDR_WHITE 203.0f

struct pq_lit {
    float m1, m2, c1, c2, c3, inv_m1, inv_m2;
    float to_pq;    // PL_COLOR_SDR_WHITE / 10000 as printed
    float from_pq;  // 10000 / PL_COLOR_SDR_WHITE as printed
};

static struct pq_lit pq_literal(void)
{
    struct pq_lit q;
    q.m1 = (float) glsl6(PQ_M1);
    q.m2 = (float) glsl6(PQ_M2);
    q.c1 = (float) glsl6(PQ_C1);
    q.c2 = (float) glsl6(PQ_C2);
    q.c3 = (float) glsl6(PQ_C3);
    q.inv_m1 = (float) (1.0 / glsl6(PQ_M1));
    q.inv_m2 = (float) (1.0 / glsl6(PQ_M2));
    q.to_pq = (float) glsl6(SDR_WHITE / 10000);
    q.from_pq = (float) glsl6(10000 / SDR_WHITE);
    return q;
}

// ------------------------------------------------------------------------------------------
// Transfer curves (pl_oracle_curve.h): built once, on first use
// ------------------------------------------------------------------------------------------

int orc_math_exact = 0;
static struct orc_curve g_curves[ORC_CV_COUNT];
static int g_curves_ready;

static void curves_init(void)
{
    if (g_curves_ready)
        return;
    #pragma omp critical(orc_curves)
    if (!g_curves_ready) {
        const struct pq_lit l = pq_literal();
        const struct orc_pq_consts lit = { l.m1, l.m2, l.c1, l.c2, l.c3, l.inv_m1, l.inv_m2 };
        const struct orc_pq_consts exact = { PQ_M1, PQ_M2, PQ_C1, PQ_C2, PQ_C3, 0, 0 };
        const double pq_pre = (float) (1.0 / glsl6(10000 / SDR_WHITE));
        const struct orc_curve_spec specs[ORC_CV_COUNT] = {
            [ORC_CV_PQ_EOTF_SDR]   = { 0, lit, l.from_pq, 0, -21, 0x3FC00000u }, // < 1.5
            [ORC_CV_PQ_EOTF]       = { 0, lit, 1.0, 0, -21, 0x3FC00000u },
            [ORC_CV_PQ_OETF_SDR]   = { 1, lit, l.to_pq, 0, -58, 0x43800000u },   // < 256
            [ORC_CV_PQ_OETF_DELIN] = { 1, lit, pq_pre, 0, -58, 0x43800000u },
            [ORC_CV_PQ_OETF]       = { 1, lit, 1.0, 0, -64, 0x41000000u },       // < 8
            [ORC_CV_PQX_OETF]      = { 1, exact, 1.0, 0, -64, 0x40000000u },     // < 2
            [ORC_CV_POW_24]        = { 2, lit, 1.0, 2.4f, -20, 0x41000000u },
            [ORC_CV_POW_INV24]     = { 2, lit, 1.0, (float) (1.0 / 2.4), -96, 0x41800000u }, // < 16
            [ORC_CV_POW_22]        = { 2, lit, 1.0, 2.2f, -20, 0x41000000u },
            [ORC_CV_POW_INV22]     = { 2, lit, 1.0, (float) (1.0 / 2.2), -96, 0x41800000u },
        };
        for (int i = 0; i < ORC_CV_COUNT; i++)
            orc_curve_build(&g_curves[i], i, &specs[i]);
        g_curves_ready = 1;
    }
}

ORC_API void orc_set_exact_math(int on) { orc_math_exact = on; }

// the table of curve `id` (tests/test_curves.py compares it with the product's)
ORC_API int orc_curve_table(int id, int *base, int *nseg, int *mbits, float *coef)
{
    if (id < 0 || id >= ORC_CV_COUNT)
        return 0;
    curves_init();
    *base = g_curves[id].base;
    *nseg = g_curves[id].nseg;
    *mbits = g_curves[id].mbits;
    if (coef)
        memcpy(coef, g_curves[id].coef, (size_t) (g_curves[id].nseg + 1) * 4 * sizeof(float));
    return 1;
}

static inline int curve(int id, float x, float *y)
{
    return !orc_math_exact && orc_curve_eval(&g_curves[id], x, y);
}

// x^g for the tabulated exponents; exact pow outside the domain
static inline float pow_cv(int id, float x, float g)
{
    float y;
    return curve(id, x, &y) ? y : pm_powf(x, g);
}

__attribute__((constructor)) static void curves_ctor(void) { curves_init(); }

// ------------------------------------------------------------------------------------------

void orc_apply_matrix(float *rgba, size_t npix, const float m[9], const float c[3])
{
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *p = rgba + 4 * i;
        float r = p[0], g = p[1], b = p[2];
        // mat3 * vec3 (+ vec3): column-major dot products; summation order x, y, z
        for (int k = 0; k < 3; k++) {
            float v = m[3*k] * r;
            v = fmaf(m[3*k + 1], g, v);
            v = fmaf(m[3*k + 2], b, v);
            p[k] = c ? v + c[k] : v;
        }
    }
}

// pl_shader_custom_lut (shaders/lut.c:212-280): [shaper_in] -> LUT -> [shaper_out]. `data`: RGB
// triples, red fastest (struct pl_custom_lut). 1-D: channel k through component k, linear
// (lut.c:700-715 form); 3-D: tetrahedral interpolation, GLSL of lut.c:773-809 statement by
// statement (six sequential conditional updates, barycentric sum in the written order).
void orc_custom_lut(float *rgba, size_t npix, const float *data, const int size[3],
                    const float *shaper_in, const float *shaper_out)
{
    if (shaper_in)
        orc_apply_matrix(rgba, npix, shaper_in, NULL);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *c = rgba + 4 * i;
        if (!size[1]) {
            for (int k = 0; k < 3; k++) {
                float fpos = clampf(c[k], 0.0f, 1.0f) * (float) (size[0] - 1);
                float fbase = floorf(fpos), fceil = ceilf(fpos);
                c[k] = mixf(data[3 * (int) fbase + k], data[3 * (int) fceil + k], fpos - fbase);
            }
            continue;
        }
        float f[3];
        int v0[3], v1[3], v2[3], v3[3];
        for (int k = 0; k < 3; k++) {
            float pos = clampf(c[k], 0.0f, 1.0f) * (float) (size[k] - 1);
            float base = floorf(pos);
            f[k] = pos - base;
            v0[k] = v1[k] = (int) base;
            v3[k] = v2[k] = (int) ceilf(pos);
        }
        // bvec3 c = greaterThanEqual(fpart.xyz, fpart.yzx)
        const bool cmp[3][3] = {
            //           x                y                z
            /* x */ { false,          f[0] >= f[1],  !(f[2] >= f[0]) },
            /* y */ { !(f[0] >= f[1]), false,         f[1] >= f[2]   },
            /* z */ { f[2] >= f[0],   !(f[1] >= f[2]), false         },
        };
        static const int idx[6][3] = { {0,1,2}, {0,2,1}, {2,0,1}, {2,1,0}, {1,2,0}, {1,0,2} };
        float s[3] = { f[0], f[1], f[2] };
        for (int t = 0; t < 6; t++) {
            const int x = idx[t][0], y = idx[t][1], z = idx[t][2];
            if (cmp[x][y] && cmp[y][z]) {
                s[0] = f[x]; s[1] = f[y]; s[2] = f[z];
                v1[x] = v3[x];
                v2[z] = v0[z];
            }
        }
#define TEXEL(v) (data + 3 * (((size_t) (v)[2] * size[1] + (v)[1]) * size[0] + (v)[0]))
        const float *t0 = TEXEL(v0), *t1 = TEXEL(v1), *t2 = TEXEL(v2), *t3 = TEXEL(v3);
#undef TEXEL
        const float w0 = 1.0f - s[0], w1 = s[0] - s[1], w2 = s[1] - s[2], w3 = s[2];
        for (int k = 0; k < 3; k++)
            c[k] = w0 * t0[k] + w1 * t1[k] + w2 * t2[k] + w3 * t3[k];
    }
    if (shaper_out)
        orc_apply_matrix(rgba, npix, shaper_out, NULL);
}

// frame mixing accumulation of pl_render_image_mix (renderer.c:3955): mix_color += w * color
void orc_mix_accumulate(float *acc, const float *rgba, size_t npix, float w)
{
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < 4 * npix; i++)
        acc[i] = fmaf(w, rgba[i], acc[i]);
}

void orc_scale(float *rgba, size_t npix, float scale, int with_alpha)
{
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        for (int k = 0; k < (with_alpha ? 4 : 3); k++)
            rgba[4*i + k] *= scale;
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_linearize — colorspace.c:589-720. trc values = enum pl_color_transfer
// ------------------------------------------------------------------------------------------

enum {
    TRC_UNKNOWN = 0, TRC_BT_1886, TRC_SRGB, TRC_LINEAR, TRC_GAMMA18, TRC_GAMMA20,
    TRC_GAMMA22, TRC_GAMMA24, TRC_GAMMA26, TRC_GAMMA28, TRC_PRO_PHOTO, TRC_ST428,
    TRC_PQ, TRC_HLG, TRC_V_LOG, TRC_S_LOG1, TRC_S_LOG2, TRC_SCRGB,
};

static const float HLG_REF = 1000.0 / SDR_WHITE;

void orc_linearize(float *rgba, size_t npix, int trc, float csp_min, float csp_max,
                   const float luma[3])
{
    if (trc == TRC_LINEAR)
        return;
    const struct pq_lit q = pq_literal();
    // BT.1886 constants, :623-626
    const float lb = powf(csp_min, 1/2.4f), lw = powf(csp_max, 1/2.4f);
    const float bt_a = powf(lw - lb, 2.4f), bt_b = lb / (lw - lb);
    // HLG, :669-670
    const float hlg_y = 1.2f * powf(1.111f, log2f(csp_max / HLG_REF));
    const float hlg_b = sqrtf(3 * powf(csp_min / csp_max, 1 / hlg_y));
    const float hlgC = (float) glsl6(HLG_C), hlgB = (float) glsl6(HLG_B);
    const float hlg_inv_a = (float) (1.0 / glsl6(HLG_A));
    const float vlogD = (float) glsl6(VLOG_D), vlogB = (float) glsl6(VLOG_B);
    const float vlog_inv_c = (float) (1.0 / glsl6(VLOG_C));
    const float slogC = (float) glsl6(SLOG_C), slogB = (float) glsl6(SLOG_B);
    const float slog_inv_a = (float) (1.0 / glsl6(SLOG_A));
    const float slogQ = (float) glsl6(SLOG_Q), slog_inv_p = (float) (1.0 / glsl6(SLOG_P));
    const float slog_inv_k2 = (float) (1.0 / glsl6(SLOG_K2));
    const float scrgb = (float) glsl6(80.0f / SDR_WHITE);
    const bool scale_out = csp_max != 1 || csp_min != 0;
    curves_init();

    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *p = rgba + 4 * i;
        float c[3] = { p[0], p[1], p[2] };
        if (trc != TRC_SCRGB) {
            for (int k = 0; k < 3; k++)
                c[k] = fmaxf(c[k], 0.0f);
        }
        bool post_scale = false;
        switch (trc) {
        case TRC_SRGB:
            for (int k = 0; k < 3; k++)
                c[k] = 0.04045f < c[k] ? pow_cv(ORC_CV_POW_24, (c[k] + 0.055f) / 1.055f, 2.4f)
                                       : c[k] * (float) (1.0 / 12.92);
            post_scale = true;
            break;
        case TRC_BT_1886:
            for (int k = 0; k < 3; k++)
                c[k] = bt_a * pow_cv(ORC_CV_POW_24, c[k] + bt_b, 2.4f);
            break;
        case TRC_GAMMA18: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], 1.8f); post_scale = true; break;
        case TRC_GAMMA20: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], 2.0f); post_scale = true; break;
        case TRC_UNKNOWN:
        case TRC_GAMMA22: for (int k = 0; k < 3; k++) c[k] = pow_cv(ORC_CV_POW_22, c[k], 2.2f); post_scale = true; break;
        case TRC_GAMMA24: for (int k = 0; k < 3; k++) c[k] = pow_cv(ORC_CV_POW_24, c[k], 2.4f); post_scale = true; break;
        case TRC_GAMMA26: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], 2.6f); post_scale = true; break;
        case TRC_GAMMA28: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], 2.8f); post_scale = true; break;
        case TRC_PRO_PHOTO:
            for (int k = 0; k < 3; k++)
                c[k] = 0.03125f < c[k] ? pm_powf(c[k], 1.8f) : c[k] * (float) (1.0 / 16.0);
            post_scale = true;
            break;
        case TRC_ST428:
            for (int k = 0; k < 3; k++)
                c[k] = (float) (52.37 / 48.0) * pm_powf(c[k], 2.6f);
            post_scale = true;
            break;
        case TRC_PQ:
            for (int k = 0; k < 3; k++) {
                float x;
                if (curve(ORC_CV_PQ_EOTF_SDR, c[k], &x)) {
                    c[k] = x;
                    continue;
                }
                x = pm_powf(c[k], q.inv_m2);
                x = fmaxf(x - q.c1, 0.0f) / (q.c2 - q.c3 * x);
                x = pm_powf(x, q.inv_m1);
                c[k] = x * q.from_pq;
            }
            break;
        case TRC_HLG: {
            for (int k = 0; k < 3; k++) {
                float x = (1 - hlg_b) * c[k] + hlg_b;
                x = 0.5f < x ? pm_expf((x - hlgC) * hlg_inv_a) + hlgB : 4.0f * x * x;
                c[k] = x * (float) (1.0 / 12.0);
            }
            float y = luma[0] * c[0];
            y = fmaf(luma[1], c[1], y);
            y = fmaf(luma[2], c[2], y);
            float g = csp_max * pm_powf(fmaxf(y, 0.0f), hlg_y - 1);
            for (int k = 0; k < 3; k++)
                c[k] *= g;
            break;
        }
        case TRC_V_LOG:
            for (int k = 0; k < 3; k++)
                c[k] = 0.181f <= c[k] ? pm_powf(10.0f, (c[k] - vlogD) * vlog_inv_c) - vlogB
                                      : (c[k] - 0.125f) * (float) (1.0 / 5.6);
            break;
        case TRC_S_LOG1:
            for (int k = 0; k < 3; k++)
                c[k] = pm_powf(10.0f, (c[k] - slogC) * slog_inv_a) - slogB;
            break;
        case TRC_S_LOG2:
            for (int k = 0; k < 3; k++)
                c[k] = slogQ <= c[k]
                     ? (pm_powf(10.0f, (c[k] - slogC) * slog_inv_a) - slogB) * slog_inv_k2
                     : (c[k] - slogQ) * slog_inv_p;
            break;
        case TRC_SCRGB:
            for (int k = 0; k < 3; k++)
                c[k] *= scrgb;
            break;
        }
        if (post_scale && scale_out) {
            for (int k = 0; k < 3; k++)
                c[k] = (csp_max - csp_min) * c[k] + csp_min;
        }
        p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_delinearize — colorspace.c:722-847
// ------------------------------------------------------------------------------------------

void orc_delinearize(float *rgba, size_t npix, int trc, float csp_min, float csp_max,
                     const float luma[3], int black_scaled)
{
    if (trc == TRC_LINEAR)
        return;
    const struct pq_lit q = pq_literal();
    const float lb = powf(csp_min, 1/2.4f), lw = powf(csp_max, 1/2.4f);
    const float bt_a = powf(lw - lb, 2.4f), bt_b = lb / (lw - lb);
    const float bt_inv_a = (float) (1.0 / bt_a);
    const float hlg_y = 1.2f * powf(1.111f, log2f(csp_max / HLG_REF));
    const float hlg_b = sqrtf(3 * powf(csp_min / csp_max, 1 / hlg_y));
    const float hlgA = (float) glsl6(HLG_A), hlgB = (float) glsl6(HLG_B), hlgC = (float) glsl6(HLG_C);
    const float vlog_c10 = (float) glsl6(VLOG_C / M_LN10), vlogB = (float) glsl6(VLOG_B),
                vlogD = (float) glsl6(VLOG_D);
    const float slog_a10 = (float) glsl6(SLOG_A / M_LN10), slogB = (float) glsl6(SLOG_B),
                slogC = (float) glsl6(SLOG_C), slogP = (float) glsl6(SLOG_P),
                slogQ = (float) glsl6(SLOG_Q), slogK2 = (float) glsl6(SLOG_K2);
    const float to_scrgb = (float) glsl6(SDR_WHITE / 80.0f);
    const float pq_pre = (float) (1.0 / glsl6(10000 / SDR_WHITE));
    const bool pre_scale = black_scaled && trc != TRC_HLG && (csp_max != 1 || csp_min != 0);
    const float ps_a = 1 / (csp_max - csp_min), ps_b = -csp_min / (csp_max - csp_min);
    curves_init();

    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *p = rgba + 4 * i;
        float c[3] = { p[0], p[1], p[2] };
        if (pre_scale) {
            for (int k = 0; k < 3; k++)
                c[k] = ps_a * c[k] + ps_b;
        }
        if (trc != TRC_SCRGB) {
            for (int k = 0; k < 3; k++)
                c[k] = fmaxf(c[k], 0.0f);
        }
        switch (trc) {
        case TRC_SRGB:
            for (int k = 0; k < 3; k++)
                c[k] = 0.0031308f <= c[k] ? 1.055f * pow_cv(ORC_CV_POW_INV24, c[k], (float) (1.0 / 2.4)) - 0.055f
                                          : c[k] * 12.92f;
            break;
        case TRC_BT_1886:
            for (int k = 0; k < 3; k++)
                c[k] = pow_cv(ORC_CV_POW_INV24, bt_inv_a * c[k], (float) (1.0 / 2.4)) - bt_b;
            break;
        case TRC_GAMMA18: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], (float) (1.0 / 1.8)); break;
        case TRC_GAMMA20: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], (float) (1.0 / 2.0)); break;
        case TRC_UNKNOWN:
        case TRC_GAMMA22: for (int k = 0; k < 3; k++) c[k] = pow_cv(ORC_CV_POW_INV22, c[k], (float) (1.0 / 2.2)); break;
        case TRC_GAMMA24: for (int k = 0; k < 3; k++) c[k] = pow_cv(ORC_CV_POW_INV24, c[k], (float) (1.0 / 2.4)); break;
        case TRC_GAMMA26: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], (float) (1.0 / 2.6)); break;
        case TRC_GAMMA28: for (int k = 0; k < 3; k++) c[k] = pm_powf(c[k], (float) (1.0 / 2.8)); break;
        case TRC_ST428:
            for (int k = 0; k < 3; k++)
                c[k] = pm_powf(c[k] * (float) (48.0 / 52.37), (float) (1.0 / 2.6));
            break;
        case TRC_PRO_PHOTO:
            for (int k = 0; k < 3; k++)
                c[k] = 0.001953f <= c[k] ? pm_powf(c[k], (float) (1.0 / 1.8)) : c[k] * 16.0f;
            break;
        case TRC_PQ:
            for (int k = 0; k < 3; k++) {
                float x;
                if (curve(ORC_CV_PQ_OETF_DELIN, c[k], &x)) {
                    c[k] = x;
                    continue;
                }
                x = c[k] * pq_pre;
                x = pm_powf(x, q.m1);
                x = (q.c1 + q.c2 * x) / (1.0f + q.c3 * x);
                c[k] = pm_powf(x, q.m2);
            }
            break;
        case TRC_HLG: {
            for (int k = 0; k < 3; k++)
                c[k] *= 1.0f / csp_max;
            float y = luma[0] * c[0];
            y = fmaf(luma[1], c[1], y);
            y = fmaf(luma[2], c[2], y);
            float g = 12.0f * pm_powf(fmaxf(1e-6f, y), (1 - hlg_y) / hlg_y);
            for (int k = 0; k < 3; k++) {
                float x = c[k] * g;
                x = 1.0f < x ? hlgA * pm_logf(x - hlgB) + hlgC : 0.5f * sqrtf(x);
                c[k] = (1 / (1 - hlg_b)) * x + (-hlg_b / (1 - hlg_b));
            }
            break;
        }
        case TRC_V_LOG:
            for (int k = 0; k < 3; k++)
                c[k] = 0.01f <= c[k] ? vlog_c10 * pm_logf(c[k] + vlogB) + vlogD
                                     : 5.6f * c[k] + 0.125f;
            break;
        case TRC_S_LOG1:
            for (int k = 0; k < 3; k++)
                c[k] = slog_a10 * pm_logf(c[k] + slogB) + slogC;
            break;
        case TRC_S_LOG2:
            for (int k = 0; k < 3; k++)
                c[k] = 0.0f <= c[k] ? slog_a10 * pm_logf(slogK2 * c[k] + slogB) + slogC
                                    : slogP * c[k] + slogQ;
            break;
        case TRC_SCRGB:
            for (int k = 0; k < 3; k++)
                c[k] *= to_scrgb;
            break;
        }
        p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
    }
}

// ------------------------------------------------------------------------------------------
// Colour-system specials of pl_shader_decode_color (:311-390) / encode_color (:474-524)
// sys: 5 = BT.2020-CL, 6 = BT.2100 PQ (ICtCp), 7 = BT.2100 HLG
// ------------------------------------------------------------------------------------------

static void mat3_mul(const float m[9], float v[3])
{
    // GLSL mat3(a,b,c, d,e,f, g,h,i) is column major; m[] holds the 9 literals in order
    float x = v[0], y = v[1], z = v[2];
    for (int k = 0; k < 3; k++) {
        float r = m[k] * x;
        r = fmaf(m[3 + k], y, r);
        r = fmaf(m[6 + k], z, r);
        v[k] = r;
    }
}

static inline float pq_eotf_lit(const struct pq_lit *q, float x)
{
    float y;
    x = fmaxf(x, 0.0f);
    if (curve(ORC_CV_PQ_EOTF, x, &y))
        return y;
    x = pm_powf(x, q->inv_m2);
    x = fmaxf(x - q->c1, 0.0f) / (q->c2 - q->c3 * x);
    return pm_powf(x, q->inv_m1);
}

static inline float pq_oetf_lit(const struct pq_lit *q, float x, int clamp0)
{
    float y;
    x = clamp0 ? fmaxf(x, 0.0f) : x;
    if (curve(ORC_CV_PQ_OETF, x, &y))
        return y;
    x = pm_powf(x, q->m1);
    x = (q->c1 + q->c2 * x) / (1.0f + q->c3 * x);
    return pm_powf(x, q->m2);
}

// Test front-end: the function behind curve `id` with its exact fallback, like
// pl_cuda_math_eval(fn = 16 + id) on the product side.
ORC_API void orc_curve_apply(int id, const float *x, float *out, size_t n)
{
    const struct pq_lit q = pq_literal();
    const float pq_pre = (float) (1.0 / glsl6(10000 / SDR_WHITE));
    for (size_t i = 0; i < n; i++) {
        float y, v = x[i];
        if (curve(id, v, &y)) {
            out[i] = y;
            continue;
        }
        switch (id) {
        case ORC_CV_PQ_EOTF_SDR:
        case ORC_CV_PQ_EOTF:
            y = pm_powf(v, q.inv_m2);
            y = fmaxf(y - q.c1, 0.0f) / (q.c2 - q.c3 * y);
            y = pm_powf(y, q.inv_m1);
            out[i] = id == ORC_CV_PQ_EOTF_SDR ? y * q.from_pq : y;
            break;
        case ORC_CV_PQ_OETF_SDR:
        case ORC_CV_PQ_OETF_DELIN:
        case ORC_CV_PQ_OETF:
            y = id == ORC_CV_PQ_OETF_SDR ? v * q.to_pq : id == ORC_CV_PQ_OETF_DELIN ? v * pq_pre : v;
            y = pm_powf(y, q.m1);
            y = (q.c1 + q.c2 * y) / (1.0f + q.c3 * y);
            out[i] = pm_powf(y, q.m2);
            break;
        case ORC_CV_PQX_OETF:
            y = pm_powf(v, PQ_M1);
            y = (PQ_C1 + PQ_C2 * y) / (1.0f + PQ_C3 * y);
            out[i] = pm_powf(y, PQ_M2);
            break;
        case ORC_CV_POW_24:    out[i] = pm_powf(v, 2.4f); break;
        case ORC_CV_POW_INV24: out[i] = pm_powf(v, (float) (1.0 / 2.4)); break;
        case ORC_CV_POW_22:    out[i] = pm_powf(v, 2.2f); break;
        case ORC_CV_POW_INV22: out[i] = pm_powf(v, (float) (1.0 / 2.2)); break;
        default:               out[i] = NAN; break;
        }
    }
}

// pl_shader_extract_features (src/shaders/colorspace.c:1383-1404), after pl_shader_linearize:
//   vec3 lms = <0.020300> * M * color.rgb;  (scalar * mat3 evaluates first: `km` = k * M)
//   lms = pow(max(lms, 0.0), vec3(M1)); lms = (c1 + c2 lms) / (1 + c3 lms); lms = pow(lms, M2);
//   float I = dot(vec3(0.400000, 0.400000, 0.200000), lms);  color = vec4(I, 0.0, 0.0, 1.0);
// `m` = the float rgb2lms matrix the host passes as a uniform (pl_ipt_rgb2lms).
void orc_extract_features(float *rgba, size_t npix, const float m[9])
{
    const struct pq_lit q = pq_literal();
    float km[9];
    for (int i = 0; i < 9; i++)
        km[i] = q.to_pq * m[i];
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *p = rgba + 4 * i;
        float lms[3];
        for (int k = 0; k < 3; k++) {
            float v = km[3*k] * p[0];
            v = fmaf(km[3*k + 1], p[1], v);
            v = fmaf(km[3*k + 2], p[2], v);
            lms[k] = pq_oetf_lit(&q, v, 1);
        }
        p[0] = fmaf(0.2f, lms[2], fmaf(0.4f, lms[1], 0.4f * lms[0]));
        p[1] = 0.0f;
        p[2] = 0.0f;
        p[3] = 1.0f;
    }
}

void orc_decode_special(float *rgba, size_t npix, int sys)
{
    const struct pq_lit q = pq_literal();
    static const float lms2rgb[9] = { 3.43661f, -0.79133f, -0.0259499f,
                                      -2.50645f, 1.98360f, -0.0989137f,
                                      0.06984f, -0.192271f, 1.12486f };
    const float hlgC = (float) glsl6(HLG_C), hlgB = (float) glsl6(HLG_B), hlgA = (float) glsl6(HLG_A);
    const float hlg_inv_a = (float) (1.0 / glsl6(HLG_A));
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *p = rgba + 4 * i;
        float c[3] = { p[0], p[1], p[2] };
        switch (sys) {
        case 5: { // BT.2020-CL, :323-341
            float g = c[1];
            c[2] = c[2] * (c[2] <= 0.0f ? 1.9404f : 1.5816f) + g;
            c[0] = c[0] * (c[0] <= 0.0f ? 1.7184f : 0.9936f) + g;
            float lin[3];
            for (int k = 0; k < 3; k++)
                lin[k] = 0.08145f <= c[k]
                       ? pm_powf((c[k] + 0.0993f) * (float) (1.0 / 1.0993), (float) (1.0 / 0.45))
                       : c[k] * (float) (1.0 / 4.5);
            float gg = (lin[1] - 0.2627f * lin[0] - 0.0593f * lin[2]) * (float) (1.0 / 0.6780);
            c[1] = 0.0181f <= gg ? 1.0993f * pm_powf(gg, 0.45f) - 0.0993f : gg * 4.5f;
            break;
        }
        case 6: // ICtCp PQ, :356-371
            for (int k = 0; k < 3; k++)
                c[k] = pq_eotf_lit(&q, c[k]);
            mat3_mul(lms2rgb, c);
            for (int k = 0; k < 3; k++)
                c[k] = pq_oetf_lit(&q, c[k], 1);
            break;
        case 7: // ICtCp HLG, :375-389
            for (int k = 0; k < 3; k++)
                c[k] = 0.5f < c[k] ? pm_expf((c[k] - hlgC) * hlg_inv_a) + hlgB : 4.0f * c[k] * c[k];
            mat3_mul(lms2rgb, c);
            for (int k = 0; k < 3; k++)
                c[k] = 1.0f < c[k] ? hlgA * pm_logf(c[k] - hlgB) + hlgC : 0.5f * sqrtf(c[k]);
            break;
        }
        p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
    }
}

void orc_encode_special(float *rgba, size_t npix, int sys)
{
    const struct pq_lit q = pq_literal();
    static const float rgb2lms[9] = { 0.412109f, 0.166748f, 0.024170f,
                                      0.523925f, 0.720459f, 0.075440f,
                                      0.063965f, 0.112793f, 0.900394f };
    const float hlgC = (float) glsl6(HLG_C), hlgB = (float) glsl6(HLG_B), hlgA = (float) glsl6(HLG_A);
    const float hlg_inv_a = (float) (1.0 / glsl6(HLG_A));
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        float *p = rgba + 4 * i;
        float c[3] = { p[0], p[1], p[2] };
        switch (sys) {
        case 5: { // :477-492
            float lin[3];
            for (int k = 0; k < 3; k++)
                lin[k] = 0.08145f <= c[k]
                       ? pm_powf((c[k] + 0.0993f) * (float) (1.0 / 1.0993), (float) (1.0 / 0.45))
                       : c[k] * (float) (1.0 / 4.5);
            float g = 0.2627f * lin[0];
            g = fmaf(0.6780f, lin[1], g);
            g = fmaf(0.0593f, lin[2], g);
            g = 0.0181f <= g ? 1.0993f * pm_powf(g, 0.45f) - 0.0993f : g * 4.5f;
            c[1] = g;
            c[2] = c[2] - g;
            c[0] = c[0] - g;
            c[2] *= c[2] <= 0.0f ? (float) (1.0 / 1.9404) : (float) (1.0 / 1.5816);
            c[0] *= c[0] <= 0.0f ? (float) (1.0 / 1.7184) : (float) (1.0 / 0.9936);
            break;
        }
        case 6: // :496-508
            for (int k = 0; k < 3; k++)
                c[k] = pq_eotf_lit(&q, c[k]);
            mat3_mul(rgb2lms, c);
            for (int k = 0; k < 3; k++)
                c[k] = pq_oetf_lit(&q, c[k], 0);
            break;
        case 7: // :512-523
            for (int k = 0; k < 3; k++)
                c[k] = 0.5f < c[k] ? pm_expf((c[k] - hlgC) * hlg_inv_a) + hlgB : 4.0f * c[k] * c[k];
            mat3_mul(rgb2lms, c);
            for (int k = 0; k < 3; k++)
                c[k] = 1.0f < c[k] ? hlgA * pm_logf(c[k] - hlgB) + hlgC : 0.5f * sqrtf(c[k]);
            break;
        }
        p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
    }
}

// ------------------------------------------------------------------------------------------
// Sigmoid — colorspace.c:851-894
// ------------------------------------------------------------------------------------------

void orc_sigmoidize(float *rgba, size_t npix, float center, float slope)
{
    const float offset = 1.0 / (1 + expf(slope * center));
    const float scale = 1.0 / (1 + expf(slope * (center - 1))) - offset;
    const float inv_slope = 1.0 / slope;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        for (int k = 0; k < 3; k++) {
            float c = clampf(rgba[4*i + k], 0.0f, 1.0f);
            rgba[4*i + k] = center - inv_slope * pm_logf(1.0f / (c * scale + offset) - 1.0f);
        }
    }
}

void orc_unsigmoidize(float *rgba, size_t npix, float center, float slope)
{
    const float offset = 1.0 / (1 + expf(slope * center));
    const float scale = 1.0 / (1 + expf(slope * (center - 1))) - offset;
    const float inv_scale = 1.0 / scale, os = offset / scale;
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < npix; i++) {
        for (int k = 0; k < 3; k++) {
            float c = clampf(rgba[4*i + k], 0.0f, 1.0f);
            rgba[4*i + k] = inv_scale / (1.0f + pm_expf(slope * (center - c))) - os;
        }
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_detect_peak — colorspace.c:1257-1350 (exact constants: ${const float:...})
// ------------------------------------------------------------------------------------------

#define PQ_BITS 14
#define PQ_MAXV ((1 << PQ_BITS) - 1)
#define HIST_BITS 7
#define HIST_BIAS (1 << (HIST_BITS - 1))
#define HIST_BINS ORC_PEAK_HIST_BINS

void orc_detect_peak(const float *rgba, int w, int h, const struct orc_peak_params *p,
                     struct orc_peak_buf *out)
{
    memset(out, 0, sizeof(*out));
    const int gw = (w + p->wg_w - 1) / p->wg_w, gh = (h + p->wg_h - 1) / p->wg_h;
    const unsigned wg_size = p->wg_w * p->wg_h;
    const float cutoff = fmaxf(p->black_cutoff, 0.0f) * 1e-2f;
    const float to_pq = SDR_WHITE / 10000.0f;

    // linearization is pointwise: do the whole image in one (parallel) call, then walk the
    // work groups serially
    float *lin = malloc((size_t) w * h * 4 * sizeof(float));
    memcpy(lin, rgba, (size_t) w * h * 4 * sizeof(float));
    orc_linearize(lin, (size_t) w * h, p->trc, p->csp_min, p->csp_max, p->luma_coeffs);

    for (int wy = 0; wy < gh; wy++) {
        for (int wx = 0; wx < gw; wx++) {
            unsigned wg_idx = wy * gw + wx, slice = wg_idx % ORC_PEAK_SLICES;
            unsigned wg_sum = 0, wg_max = 0, wg_black = 0;
            unsigned wg_hist[HIST_BINS] = {0};
            for (int ly = 0; ly < p->wg_h; ly++) {
                for (int lx = 0; lx < p->wg_w; lx++) {
                    int x = wx * p->wg_w + lx, y = wy * p->wg_h + ly;
                    x = x < w ? x : w - 1; // clamped texture fetch beyond the edge
                    y = y < h ? y : h - 1;
                    const float *px = lin + ((size_t) y * w + x) * 4;
                    float luma = p->luma_coeffs[0] * px[0];
                    luma = fmaf(p->luma_coeffs[1], px[1], luma);
                    luma = fmaf(p->luma_coeffs[2], px[2], luma);
                    luma *= to_pq;
                    luma = clampf(luma, 0.0f, 1.0f);
                    float pqv;
                    if (curve(ORC_CV_PQX_OETF, luma, &pqv)) {
                        luma = pqv;
                    } else {
                        luma = pm_powf(luma, PQ_M1);
                        luma = (PQ_C1 + PQ_C2 * luma) / (1.0f + PQ_C3 * luma);
                        luma = pm_powf(luma, PQ_M2);
                    }
                    if (cutoff)
                        luma *= smoothstepf(0.0f, cutoff, luma);
                    unsigned y_pq = (unsigned) ((float) PQ_MAXV * luma);
                    if (p->use_histogram) {
                        int bin = (int) y_pq >> (PQ_BITS - HIST_BITS);
                        bin -= HIST_BIAS;
                        bin = bin < 0 ? 0 : bin > HIST_BINS - 1 ? HIST_BINS - 1 : bin;
                        wg_hist[bin]++;
                    }
                    wg_sum += y_pq;
                    wg_max = y_pq > wg_max ? y_pq : wg_max;
                    if (cutoff && y_pq == 0)
                        wg_black++;
                }
            }
            if (p->use_histogram) {
                if (cutoff)
                    wg_hist[0] -= wg_black;
                for (int i = 0; i < HIST_BINS; i++)
                    out->frame_hist[slice][i] += wg_hist[i];
            }
            unsigned num = wg_size - wg_black;
            out->frame_wg_count[slice] += 1;
            out->frame_wg_active[slice] += num < 1 ? num : 1;
            if (num > 0) {
                out->frame_sum_pq[slice] += wg_sum / num;
                if (wg_max > out->frame_max_pq[slice])
                    out->frame_max_pq[slice] = wg_max;
            }
        }
    }
    free(lin);
}

// measure_peak, colorspace.c:1020-1069
#define HIST_PQ(bin) (((bin) + HIST_BIAS) << (PQ_BITS - HIST_BITS))

static float measure_peak(const struct orc_peak_buf *data, float percentile)
{
    unsigned frame_max_pq = data->frame_max_pq[0];
    for (int k = 1; k < ORC_PEAK_SLICES; k++)
        frame_max_pq = frame_max_pq > data->frame_max_pq[k] ? frame_max_pq : data->frame_max_pq[k];
    const float frame_max = (float) frame_max_pq / PQ_MAXV;
    if (percentile <= 0 || percentile >= 100)
        return frame_max;
    unsigned total_pixels = 0;
    for (int k = 0; k < ORC_PEAK_SLICES; k++)
        for (int i = 0; i < HIST_BINS; i++)
            total_pixels += data->frame_hist[k][i];
    if (!total_pixels)
        return frame_max;
    const unsigned target_pixel = ceilf(percentile / 100.0f * total_pixels);
    if (target_pixel >= total_pixels)
        return frame_max;
    unsigned sum = 0;
    for (int i = 0; i < HIST_BINS; i++) {
        unsigned next = sum;
        for (int k = 0; k < ORC_PEAK_SLICES; k++)
            next += data->frame_hist[k][i];
        if (next < target_pixel) {
            sum = next;
            continue;
        }
        const unsigned count_low = sum, count_high = next + 1;
        const float pq_low = (float) HIST_PQ(i) / PQ_MAXV;
        float pq_high = (float) HIST_PQ(i + 1) / PQ_MAXV;
        if (count_high > total_pixels)
            pq_high = frame_max;
        const float ratio = (float) (target_pixel - count_low) / (count_high - count_low);
        return (1 - ratio) * pq_low + ratio * pq_high; // PL_MIX
    }
    return frame_max;
}

// update_peak_buf, colorspace.c:1072-1153
int orc_peak_update(struct orc_peak_state *st, const struct orc_peak_buf *data,
                    float smoothing_period, float scene_low, float scene_high, float percentile)
{
    if (!(data->frame_wg_count[0] > 0))
        return 0;
    uint64_t frame_sum_pq = 0, frame_wg_count = 0, frame_wg_active = 0;
    for (int k = 0; k < ORC_PEAK_SLICES; k++) {
        frame_sum_pq += data->frame_sum_pq[k];
        frame_wg_count += data->frame_wg_count[k];
        frame_wg_active += data->frame_wg_active[k];
    }
    float avg_pq, max_pq;
    if (frame_wg_active) {
        avg_pq = (float) frame_sum_pq / (frame_wg_active * PQ_MAXV);
        max_pq = measure_peak(data, percentile);
    } else {
        avg_pq = max_pq = 1e-6f; // PL_COLOR_HDR_BLACK
    }
    if (!st->avg_pq) {
        st->avg_pq = avg_pq;
        st->max_pq = max_pq;
    } else {
        const float epsilon = 1.0f / PQ_MAXV;
        if (fabsf(avg_pq - st->avg_pq) < epsilon)
            avg_pq = st->avg_pq;
        if (fabsf(max_pq - st->max_pq) < epsilon)
            max_pq = st->max_pq;
    }
    const float coeff = smoothing_period ? 1.0f - expf(-1.0f / smoothing_period) : 1.0f;
    st->avg_pq += coeff * (avg_pq - st->avg_pq);
    st->max_pq += coeff * (max_pq - st->max_pq);
    if (scene_low > 0 && scene_high > 0) {
        const float log10_pq = 1e-2f;
        const float thresh_low = scene_low * log10_pq, thresh_high = scene_high * log10_pq;
        const float bias = (float) frame_wg_active / frame_wg_count;
        const float delta = bias * fabsf(avg_pq - st->avg_pq);
        // pl_smoothstep (src/common.h)
        float t = (delta - thresh_low) / (thresh_high - thresh_low);
        t = t < 0 ? 0 : t > 1 ? 1 : t;
        const float mix_coeff = t * t * (3.0f - 2.0f * t);
        st->avg_pq = (1 - mix_coeff) * st->avg_pq + mix_coeff * avg_pq;
        st->max_pq = (1 - mix_coeff) * st->max_pq + mix_coeff * max_pq;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------
// pl_shader_color_map_ex — colorspace.c:1612-2024 (GLSL :1764-2022)
// ------------------------------------------------------------------------------------------

static void matv(const float m[9], const float v[3], float o[3])
{
    // SH_MAT3 uploads the row-major pl_matrix3x3 transposed: result = M * v
    for (int k = 0; k < 3; k++) {
        float r = m[3*k] * v[0];
        r = fmaf(m[3*k + 1], v[1], r);
        r = fmaf(m[3*k + 2], v[2], r);
        o[k] = r;
    }
}

// pl_ipt_lms2ipt / pl_ipt_ipt2lms: Ebner & Fairchild IPT (1998) and its published numeric
// inverse, as tabulated in src/colorspace.c:1377-1390. Part of the shader (SH_MAT3 constants,
// colorspace.c:1769-1770). tests/test_oracle_kat.py pins both against the reference library.
static const float LMS2IPT[9] = {
    0.4000f, 0.4000f, 0.2000f,
    4.4550f, -4.8510f, 0.3960f,
    0.8056f, 0.3572f, -1.1628f,
};
static const float IPT2LMS[9] = {
    1.0f, 0.0975689f, 0.205226f,
    1.0f, -0.1138760f, 0.133217f,
    1.0f, 0.0326151f, -0.676887f,
};

ORC_API void orc_get_ipt_matrices(float lms2ipt[9], float ipt2lms[9])
{
    memcpy(lms2ipt, LMS2IPT, sizeof(LMS2IPT));
    memcpy(ipt2lms, IPT2LMS, sizeof(IPT2LMS));
}

static inline const float *gamut_texel(const struct orc_color_map_params *p, int i, int c, int h)
{
    return p->gamut_lut + (((size_t) h * p->lut_C + c) * p->lut_I + i) * 4;
}

static inline float satf(float x) { return x > 0.0f ? (x < 1.0f ? x : 1.0f) : 0.0f; } // NaN -> 0
static inline float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }

// Trilinear 3DLUT fetch. The reference binds the LUT as a linearly filtered 3D texture and
// addresses texel centres through LUT_POS (lut.c:285-295: x -> mix(0.5/n, 1 - 0.5/n, x)), which
// in texel space is x (n - 1); the hardware filter clamps to the edge and interpolates with
// 8-bit weights. Restated in fp32: saturate the index, t = x (n - 1), lerp(a, b, t) =
// fma(t, b - a, a), axes in the order I, C, h.
static void gamut_sample(const struct orc_color_map_params *p, const float idx[3], float out[3])
{
    const int n[3] = { p->lut_I, p->lut_C, p->lut_h };
    int i0[3], i1[3];
    float f[3];
    for (int a = 0; a < 3; a++) {
        const float t = satf(idx[a]) * (float) (n[a] - 1);
        const float fl = floorf(t);
        i0[a] = (int) fl;
        i1[a] = i0[a] + 1 < n[a] - 1 ? i0[a] + 1 : n[a] - 1;
        f[a] = t - fl;
    }
    const float *t000 = gamut_texel(p, i0[0], i0[1], i0[2]), *t001 = gamut_texel(p, i1[0], i0[1], i0[2]);
    const float *t010 = gamut_texel(p, i0[0], i1[1], i0[2]), *t011 = gamut_texel(p, i1[0], i1[1], i0[2]);
    const float *t100 = gamut_texel(p, i0[0], i0[1], i1[2]), *t101 = gamut_texel(p, i1[0], i0[1], i1[2]);
    const float *t110 = gamut_texel(p, i0[0], i1[1], i1[2]), *t111 = gamut_texel(p, i1[0], i1[1], i1[2]);
    for (int k = 0; k < 3; k++) {
        const float x00 = lerpf(t000[k], t001[k], f[0]), x01 = lerpf(t010[k], t011[k], f[0]);
        const float x10 = lerpf(t100[k], t101[k], f[0]), x11 = lerpf(t110[k], t111[k], f[0]);
        const float y0 = lerpf(x00, x01, f[1]), y1 = lerpf(x10, x11, f[1]);
        out[k] = lerpf(y0, y1, f[2]);
    }
}

static inline float lut1d_linear(const float *lut, int n, float x)
{
    float fpos = clampf(x, 0.0f, 1.0f) * (float) (n - 1);
    float fbase = floorf(fpos), fceil = ceilf(fpos);
    return mixf(lut[(int) fbase], lut[(int) fceil], fpos - fbase);
}

// `tone_map(x)` of pl_shader_color_map_ex: clip / linear fast paths or the 1-D LUT (:1826-1876)
static inline float tone_map_x(const struct orc_color_map_params *p, float x, float lin_a,
                               float lin_b, float tone_a, float tone_b)
{
    if (p->tone_kind == 1)
        return clampf(x, p->tone_input_min, p->tone_input_max);
    if (p->tone_kind == 2) {
        x = lin_a * x + lin_b;
        x = clampf(x, 0.0f, 1.0f);
        return (p->tone_output_max - p->tone_output_min) * x + p->tone_output_min;
    }
    return lut1d_linear(p->tone_lut, p->tone_lut_size, tone_a * x + tone_b);
}

void orc_color_map(float *rgba, size_t npix, const struct orc_color_map_params *p)
{
    const struct pq_lit q = pq_literal();
    if (!p->prelinearized)
        orc_linearize(rgba, npix, p->src_trc, p->src_min, p->src_max, p->src_luma);

    if (!p->need_tone_map && !p->need_gamut_map) {
        // fast path :1782-1789: lms2rgb already holds lms2rgb*rgb2lms (or identity -> skipped
        // by the caller passing the identity matrix)
        orc_apply_matrix(rgba, npix, p->lms2rgb, NULL);
        orc_delinearize(rgba, npix, p->dst_trc, p->dst_min, p->dst_max, p->dst_luma,
                        p->dst_black_scaled);
        return;
    }

    const float lut_range = p->tone_input_max - p->tone_input_min;
    const float tone_a = 1.0f / lut_range, tone_b = -p->tone_input_min / lut_range;
    const float lin_scale = p->tone_input_max - p->tone_input_min;
    const float lin_a = p->tone_exposure / lin_scale;
    const float lin_b = -p->tone_exposure / lin_scale * p->tone_input_min;
    const float g_range = p->gamut_max_luma - p->gamut_min_luma;
    const float g_a = 1.0f / g_range, g_b = -p->gamut_min_luma / g_range;
    const float inv_2pi = (float) glsl6(0.5f / M_PI);
    const float chroma_bias = (float) (32768.0 / 65535.0);
    curves_init();

    #pragma omp parallel for schedule(static)
    for (size_t n = 0; n < npix; n++) {
        float *px = rgba + 4 * n;
        float lms[3], lmspq[3], ipt[3];
        matv(p->rgb2lms, px, lms);
        for (int k = 0; k < 3; k++) {
            float x = fmaxf(lms[k], 0.0f);
            if (curve(ORC_CV_PQ_OETF_SDR, x, &lmspq[k]))
                continue;
            x = pm_powf(x * q.to_pq, q.m1);
            x = (q.c1 + q.c2 * x) / (1.0f + q.c3 * x);
            lmspq[k] = pm_powf(x, q.m2);
        }
        matv(LMS2IPT, lmspq, ipt);
        float i_orig = ipt[0];

        if (p->need_tone_map) {
            float x = ipt[0];
            // contrast recovery :1879-1921 (only when tone.input_max >= tone.output_max; the
            // caller passes feature_map = NULL otherwise)
            if (p->feature_map && p->contrast_recovery) {
                const struct orc_tex *fm = p->feature_map;
                // `pos` of sh_bind without a rect: the whole feature map over the output
                const float lpos[2] = { (float) (((double) (n % p->width) + 0.5) / p->width),
                                        (float) (((double) (n / p->width) + 0.5) / p->height) };
                const float lpt[2] = { (float) (1.0 / fm->w), (float) (1.0 / fm->h) };
                const float lsize[2] = { (float) fm->w, (float) fm->h };
                float g[4], h[4];
                for (int a = 0; a < 2; a++) {
                    const float t = lpos[a] * lsize[a] + 0.5f;
                    const float frac = t - floorf(t), frac2 = frac * frac;
                    const float inv = 1.0f - frac, inv2 = inv * inv;
                    const float w0 = (1.0f / 6.0f) * inv2 * inv;
                    const float w1 = (2.0f / 3.0f) - 0.5f * frac2 * (2.0f - frac);
                    const float w2 = (2.0f / 3.0f) - 0.5f * inv2 * (2.0f - inv);
                    const float w3 = (1.0f / 6.0f) * frac2 * frac;
                    g[a] = w0 + w1;
                    g[2 + a] = w2 + w3;
                    h[a] = w1 / g[a] + inv;
                    h[2 + a] = w3 / g[2 + a] + inv;
                    h[a] -= 2.0f;
                }
                const float px_ = lpos[0] + lpt[0] * h[0], py_ = lpos[1] + lpt[1] * h[1];
                const float pz_ = lpos[0] + lpt[0] * h[2], pw_ = lpos[1] + lpt[1] * h[3];
                const float l00 = orc_tex_linear_r(fm, px_, py_), l01 = orc_tex_linear_r(fm, px_, pw_);
                const float l0 = l01 * (1.0f - g[1]) + l00 * g[1];
                const float l10 = orc_tex_linear_r(fm, pz_, py_), l11 = orc_tex_linear_r(fm, pz_, pw_);
                const float l1 = l11 * (1.0f - g[1]) + l10 * g[1];
                const float luma = l1 * (1.0f - g[0]) + l0 * g[0];
                const float highres = clampf(ipt[0], 0.0f, 1.0f);
                const float lowres = clampf(luma, 0.0f, 1.0f);
                const float detail = highres - lowres;
                const float base = tone_map_x(p, highres, lin_a, lin_b, tone_a, tone_b);
                const float sharp = tone_map_x(p, lowres, lin_a, lin_b, tone_a, tone_b) + detail;
                x = base * (1.0f - p->contrast_recovery) + sharp * p->contrast_recovery;
                x = clampf(x, p->tone_output_min, p->tone_output_max);
            } else {
                x = tone_map_x(p, x, lin_a, lin_b, tone_a, tone_b);
            }
            ipt[0] = x;
            // hull desaturation :1930-1932
            float h0 = ((i_orig - 6.0f) * i_orig + 9.0f) * i_orig;
            float h1 = ((ipt[0] - 6.0f) * ipt[0] + 9.0f) * ipt[0];
            float s = fminf(i_orig / ipt[0], h1 / h0);
            ipt[1] *= s;
            ipt[2] *= s;
        }

        if (p->need_gamut_map) {
            float idx[3];
            idx[0] = g_a * ipt[0] + g_b;
            idx[1] = 2.0f * sqrtf(fmaf(ipt[1], ipt[1], ipt[2] * ipt[2]));
            idx[2] = inv_2pi * pf_atan2f(ipt[2], ipt[1]) + 0.5f;
            gamut_sample(p, idx, ipt);
            ipt[1] -= chroma_bias;
            ipt[2] -= chroma_bias;
        }

        matv(IPT2LMS, ipt, lmspq);
        for (int k = 0; k < 3; k++) {
            float x = fmaxf(lmspq[k], 0.0f);
            if (curve(ORC_CV_PQ_EOTF_SDR, x, &lms[k]))
                continue;
            x = pm_powf(x, q.inv_m2);
            x = fmaxf(x - q.c1, 0.0f) / (q.c2 - q.c3 * x);
            x = pm_powf(x, q.inv_m1);
            lms[k] = x * q.from_pq;
        }
        matv(p->lms2rgb, lms, px);
    }

    orc_delinearize(rgba, npix, p->dst_trc, p->dst_min, p->dst_max, p->dst_luma,
                    p->dst_black_scaled);
}

// ------------------------------------------------------------------------------------------
// pl_shader_dither — dithering.c:109-274
// ------------------------------------------------------------------------------------------

void orc_prng(uint32_t state[3], float out[3]);

void orc_dither(float *rgba, int w, int h, const struct orc_dither_params *p)
{
    int size = 0;
    if (p->method == 0 || p->method == 1)
        size = p->lut_size;
    else if (p->method == 2)
        size = 16;
    const float inv_size = size ? 1.0f / (float) size : 0.0f; // `1.0/$size` folded
    const float scale = (float) ((1ull << p->new_depth) - 1);
    const float inv_scale = 1.0f / scale;

    float rot[2][2] = {{1, 0}, {0, 1}};
    if (p->temporal) {
        int phase = p->frame_index % 8;
        float r = phase * (M_PI / 2);
        float m = phase < 4 ? 1 : -1;
        rot[0][0] = cos(r);     rot[0][1] = -sin(r);
        rot[1][0] = sin(r) * m; rot[1][1] = cos(r) * m;
    }

    #pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float *c = rgba + ((size_t) y * w + x) * 4;
            float bias = 0.0f;
            float px = 0, py = 0;
            if (size) {
                px = ((float) x + 0.5f) * inv_size;
                py = ((float) y + 0.5f) * inv_size;
                px -= floorf(px);
                py -= floorf(py);
                if (p->temporal) {
                    // pos = fract(rot * pos + 1.0); mat2 uploaded from row-major data as
                    // column-major: rot*pos = (m00*x + m10*y, m01*x + m11*y)
                    float nx = fmaf(rot[1][0], py, rot[0][0] * px) + 1.0f;
                    float ny = fmaf(rot[1][1], py, rot[0][1] * px) + 1.0f;
                    px = nx - floorf(nx);
                    py = ny - floorf(ny);
                }
            }
            switch (p->method) {
            case 3: {
                uint32_t st[3] = { (uint32_t) x, (uint32_t) y, p->temporal ? p->frame_index : 0 };
                float rnd[3];
                orc_prng(st, rnd);
                bias = rnd[0];
                break;
            }
            case 2: {
                uint32_t xx = (uint32_t) (px * 16.0f) % 16u, yy = (uint32_t) (py * 16.0f) % 16u;
                xx = xx ^ yy;
                xx = (xx | xx << 2) & 0x33333333u; yy = (yy | yy << 2) & 0x33333333u;
                xx = (xx | xx << 1) & 0x55555555u; yy = (yy | yy << 1) & 0x55555555u;
                uint32_t b = xx + (yy << 1);
                b = (b * 0x0802u & 0x22110u) | (b * 0x8020u & 0x88440u);
                b = 0x10101u * b;
                b = (b >> 16) & 0xFFu;
                bias = (float) b * (1.0f / 256.0f);
                break;
            }
            default: {
                int ix = (int) (px * (float) size), iy = (int) (py * (float) size);
                bias = p->lut[iy * size + ix];
                break;
            }
            }

            if (p->gamma != 1.0f && p->new_depth <= 4) {
                for (int k = 0; k < 4; k++) {
                    float col = c[k];
                    float col_lin = pm_powf(col, p->gamma);
                    float low, high, offset;
                    if (p->new_depth == 1) {
                        low = 0.0f; high = 1.0f; offset = col_lin;
                    } else {
                        low = floorf(col * scale) / scale;
                        high = ceilf(col * scale) / scale;
                        float low_lin = pm_powf(low, p->gamma), high_lin = pm_powf(high, p->gamma);
                        float range = high_lin - low_lin;
                        offset = (col_lin - low_lin) / fmaxf(range, 1e-6f);
                    }
                    c[k] = offset > bias ? high : low;
                }
            } else {
                for (int k = 0; k < 4; k++) {
                    float v = scale * c[k] + bias;
                    c[k] = floorf(v) * inv_scale;
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// pl_shader_dovi_reshape — colorspace.c:106-273 (+ reshape_mmr :52-92, reshape_poly :94-97)
// ------------------------------------------------------------------------------------------
static inline float dot3f(const float a[3], const float b[3])
{
    return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}

static inline float dot4f(const float a[4], const float b[4])
{
    return fmaf(a[3], b[3], fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])));
}

void orc_dovi_reshape(float *rgba, size_t npix, const struct orc_dovi_comp comp[3])
{
    // host-side packing exactly like :128-169
    float coeffs[3][8][4], mmr[3][48][4], pivots[3][7];
    memset(coeffs, 0, sizeof(coeffs));
    memset(mmr, 0, sizeof(mmr));
    for (int c = 0; c < 3; c++) {
        const struct orc_dovi_comp *cp = &comp[c];
        int mmr_idx = 0;
        for (int i = 0; i < cp->num_pivots - 1; i++) {
            if (cp->method[i] == 0) {
                for (int k = 0; k < 3; k++)
                    coeffs[c][i][k] = cp->poly_coeffs[i][k];
            } else {
                coeffs[c][i][3] = (float) cp->mmr_order[i];
                coeffs[c][i][0] = cp->mmr_constant[i];
                coeffs[c][i][1] = (float) mmr_idx;
                for (int j = 0; j < cp->mmr_order[i]; j++) {
                    float *m = &mmr[c][mmr_idx][0];
                    m[0] = cp->mmr_coeffs[i][j][0]; m[1] = cp->mmr_coeffs[i][j][1];
                    m[2] = cp->mmr_coeffs[i][j][2]; m[3] = 0.0f;
                    m[4] = cp->mmr_coeffs[i][j][3]; m[5] = cp->mmr_coeffs[i][j][4];
                    m[6] = cp->mmr_coeffs[i][j][5]; m[7] = cp->mmr_coeffs[i][j][6];
                    mmr_idx += 2;
                }
            }
        }
        for (int i = 0; i < 7; i++)
            pivots[c][i] = i < cp->num_pivots - 2 ? cp->pivots[i + 1] : 1e9f;
    }

    #pragma omp parallel for schedule(static)
    for (size_t n = 0; n < npix; n++) {
        float *px = rgba + 4 * n;
        const float sig[3] = { clampf(px[0], 0.0f, 1.0f), clampf(px[1], 0.0f, 1.0f),
                               clampf(px[2], 0.0f, 1.0f) };
        for (int c = 0; c < 3; c++) {
            const struct orc_dovi_comp *cp = &comp[c];
            if (!cp->num_pivots)
                continue;
            float s = sig[c];
            const float *co;
            if (cp->num_pivots > 2) {
                // mix tree :206-216: mix(a, b, bvec) picks b where true
                #define TESTP(i) (s >= pivots[c][i])
                int k;
                if (TESTP(3))
                    k = TESTP(5) ? (TESTP(6) ? 7 : 6) : (TESTP(4) ? 5 : 4);
                else
                    k = TESTP(1) ? (TESTP(2) ? 3 : 2) : (TESTP(0) ? 1 : 0);
                #undef TESTP
                co = coeffs[c][k];
            } else {
                co = coeffs[c][0];
            }
            if (co[3] == 0.0f) {
                s = (co[2] * s + co[1]) * s + co[0];
            } else {
                const int idx = (int) co[1];
                const unsigned order = (unsigned) co[3];
                float sigX[4] = { sig[0] * sig[1], sig[0] * sig[2], sig[1] * sig[2], 0.0f };
                sigX[3] = sigX[0] * sig[2];
                s = co[0];
                s += dot3f(mmr[c][idx + 0], sig);
                s += dot4f(mmr[c][idx + 1], sigX);
                if (order >= 2) {
                    const float sig2[3] = { sig[0] * sig[0], sig[1] * sig[1], sig[2] * sig[2] };
                    const float sigX2[4] = { sigX[0] * sigX[0], sigX[1] * sigX[1],
                                             sigX[2] * sigX[2], sigX[3] * sigX[3] };
                    s += dot3f(mmr[c][idx + 2], sig2);
                    s += dot4f(mmr[c][idx + 3], sigX2);
                    if (order >= 3) {
                        const float sig3[3] = { sig2[0] * sig[0], sig2[1] * sig[1], sig2[2] * sig[2] };
                        const float sigX3[4] = { sigX2[0] * sigX[0], sigX2[1] * sigX[1],
                                                 sigX2[2] * sigX[2], sigX2[3] * sigX[3] };
                        s += dot3f(mmr[c][idx + 4], sig3);
                        s += dot4f(mmr[c][idx + 5], sigX3);
                    }
                }
            }
            px[c] = clampf(s, cp->pivots[0], cp->pivots[cp->num_pivots - 1]);
        }
    }
}

void orc_dovi_lms(float *rgba, size_t npix, const float lms2rgb[9])
{
    const struct pq_lit q = pq_literal();
    #pragma omp parallel for schedule(static)
    for (size_t n = 0; n < npix; n++) {
        float *px = rgba + 4 * n;
        float c[3];
        for (int k = 0; k < 3; k++)
            c[k] = pq_eotf_lit(&q, px[k]);
        // `lms2rgb` is the row-major pl_matrix3x3 (uploaded transposed, i.e. M * v as written)
        float o[3];
        for (int k = 0; k < 3; k++) {
            float v = lms2rgb[3*k] * c[0];
            v = fmaf(lms2rgb[3*k + 1], c[1], v);
            o[k] = fmaf(lms2rgb[3*k + 2], c[2], v);
        }
        for (int k = 0; k < 3; k++)
            px[k] = pq_oetf_lit(&q, o[k], true);
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_error_diffusion — dithering.c:326-527. Serial restatement in shifted-column order
// with the same packed RGB8 ring words (the packing's wrap-around is part of the semantics).
// ------------------------------------------------------------------------------------------
void orc_error_diffusion(const float *rgba_in, float *rgba_out, int w, int h,
                         int new_depth, int shift, int divisor, const int *pattern)
{
    const int ring_rows = h + 2;
    int rightmost = 0;
    for (int dy = 0; dy <= 2; dy++) {
        for (int dx = -2; dx <= 2; dx++) {
            if (pattern[dy * 5 + dx + 2] && dx + dy * shift > rightmost)
                rightmost = dx + dy * shift;
        }
    }
    const uint32_t ring_size = (uint32_t) ring_rows * (uint32_t) (rightmost + 1);
    uint32_t *err = calloc(ring_size, sizeof(uint32_t));
    const float quant = (float) ((1ll << new_depth) - 1);
    const int shifted_width = w + (h - 1) * shift;
    for (int xs = 0; xs < shifted_width; xs++) {
        for (int y = 0; y < h; y++) {
            const int x = xs - y * shift;
            if (x < 0 || x >= w)
                continue;
            const uint32_t idx = (uint32_t) (xs * ring_rows + y) % ring_size;
            const float *po = rgba_in + ((size_t) y * w + x) * 4;
            float *out = rgba_out + ((size_t) y * w + x) * 4;
            uint32_t e = err[idx] + ((128u << 24) | (128u << 12) | 128u);
            err[idx] = 0;
            const int ei[3] = { (int) ((e >> 24) & 0xFFu) - 128, (int) ((e >> 12) & 0xFFu) - 128,
                                (int) (e & 0xFFu) - 128 };
            float pix[3], dith[3], ed[3];
            for (int k = 0; k < 3; k++) {
                pix[k] = po[k] * quant + (float) ei[k] / 254.0f;
                dith[k] = roundf(pix[k]);
                out[k] = dith[k] / quant;
                ed[k] = (pix[k] - dith[k]) * 254.0f / (float) divisor;
            }
            out[3] = po[3];
            for (int dividend = 1; dividend <= divisor; dividend++) {
                bool assigned = false;
                uint32_t packed = 0;
                for (int dy = 0; dy <= 2; dy++) {
                    for (int dx = -2; dx <= 2; dx++) {
                        if (pattern[dy * 5 + dx + 2] != dividend)
                            continue;
                        if (!assigned) {
                            assigned = true;
                            const int t0 = (int) roundf(ed[0] * (float) dividend);
                            const int t1 = (int) roundf(ed[1] * (float) dividend);
                            const int t2 = (int) roundf(ed[2] * (float) dividend);
                            packed = ((uint32_t) (t0 & 0xFF) << 24) | ((uint32_t) (t1 & 0xFF) << 12) |
                                     (uint32_t) (t2 & 0xFF);
                        }
                        if (dx < 0 && !(x >= -dx))
                            continue;
                        const uint32_t delta = (uint32_t) ((dx + dy * shift) * ring_rows + dy);
                        err[(idx + delta) % ring_size] += packed;
                    }
                }
            }
        }
    }
    free(err);
}
