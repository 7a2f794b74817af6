// Per-pixel colour op interpreter: executes the typed op list a pl_shader recorded, in order.
// Each op is the fp32 restatement of one GLSL snippet of the reference:
//   MATRIX           src/shaders/colorspace.c:308, :569
//   LINEARIZE        pl_shader_linearize    colorspace.c:589-720
//   DELINEARIZE      pl_shader_delinearize  colorspace.c:722-847
//   DE/ENCODE_SPECIAL ICtCp & BT.2020-CL    colorspace.c:311-390, :474-524
//   SIGMOID          colorspace.c:851-894
//   COLOR_MAP        pl_shader_color_map_ex colorspace.c:1791-1995
//   DITHER           pl_shader_dither       dithering.c:181-271
// The switch is warp-uniform (the program is a kernel parameter), so there is no divergence.
#pragma once

#include "kernels_common.cuh"
#include "plb_math.cuh"

enum {
    TRC_UNKNOWN = 0, TRC_BT_1886, TRC_SRGB, TRC_LINEAR, TRC_GAMMA18, TRC_GAMMA20,
    TRC_GAMMA22, TRC_GAMMA24, TRC_GAMMA26, TRC_GAMMA28, TRC_PRO_PHOTO, TRC_ST428,
    TRC_PQ, TRC_HLG, TRC_V_LOG, TRC_S_LOG1, TRC_S_LOG2, TRC_SCRGB,
};

PLB_DEV float plb_dot3_fma(float a0, float a1, float a2, float x, float y, float z)
{
    float r = a0 * x;
    r = __fmaf_rn(a1, y, r);
    r = __fmaf_rn(a2, z, r);
    return r;
}

// M (row-major) * v
PLB_DEV void plb_matv(const float *m, float &r, float &g, float &b)
{
    float x = r, y = g, z = b;
    r = plb_dot3_fma(m[0], m[1], m[2], x, y, z);
    g = plb_dot3_fma(m[3], m[4], m[5], x, y, z);
    b = plb_dot3_fma(m[6], m[7], m[8], x, y, z);
}

// GLSL mat3(9 literals) * v, column-major literal order
PLB_DEV void plb_mat3_lit(const float *m, float &r, float &g, float &b)
{
    float x = r, y = g, z = b;
    r = plb_dot3_fma(m[0], m[3], m[6], x, y, z);
    g = plb_dot3_fma(m[1], m[4], m[7], x, y, z);
    b = plb_dot3_fma(m[2], m[5], m[8], x, y, z);
}

// PQ EOTF / OETF (colorspace.c:662-668, :778-784). The composite function is a transfer curve
// (curves.h); arguments outside its domain take the exact pow / divide / pow chain, out of line.
static __device__ __noinline__ float plb_pq_eotf_slow(float x)
{
    x = plm_powf(x, c_pq.inv_m2);
    x = fmaxf(x - c_pq.c1, 0.0f) / (c_pq.c2 - c_pq.c3 * x);
    return plm_powf(x, c_pq.inv_m1);
}

static __device__ __noinline__ float plb_pq_oetf_slow(float x)
{
    x = plm_powf(x, c_pq.m1);
    x = (c_pq.c1 + c_pq.c2 * x) / (1.0f + c_pq.c3 * x);
    return plm_powf(x, c_pq.m2);
}

// The exact chain behind curve CV (what an argument outside the tabulated domain evaluates to)
template <int CV>
static __device__ __noinline__ float plb_curve_exact(float x)
{
    if (CV == PLB_CV_PQ_EOTF_SDR)
        return plb_pq_eotf_slow(x) * c_pq.from_pq;
    if (CV == PLB_CV_PQ_EOTF)
        return plb_pq_eotf_slow(x);
    if (CV == PLB_CV_PQ_OETF_SDR)
        return plb_pq_oetf_slow(x * c_pq.to_pq);
    if (CV == PLB_CV_PQ_OETF_DELIN)
        return plb_pq_oetf_slow(x * c_pq.inv_from_pq);
    if (CV == PLB_CV_PQ_OETF)
        return plb_pq_oetf_slow(x);
    if (CV == PLB_CV_PQX_OETF) { // exact float constants (colorspace.c:1279-1288)
        const float PQ_M1 = 2610.f / 4096 * 1.f / 4, PQ_M2 = 2523.f / 4096 * 128,
                    PQ_C1 = 3424.f / 4096, PQ_C2 = 2413.f / 4096 * 32, PQ_C3 = 2392.f / 4096 * 32;
        x = plm_powf(x, PQ_M1);
        x = (PQ_C1 + PQ_C2 * x) / (1.0f + PQ_C3 * x);
        return plm_powf(x, PQ_M2);
    }
    const float g = CV == PLB_CV_POW_24 ? 2.4f : CV == PLB_CV_POW_INV24 ? (float) (1.0 / 2.4)
                  : CV == PLB_CV_POW_22 ? 2.2f : (float) (1.0 / 2.2);
    return plm_powf(x, g);
}

// General kernels: descriptor-driven curve (honours exact-math mode), exact chain on a miss
template <int CV>
PLB_DEV float plb_curve(float x)
{
    float y;
    return plm_curve<CV>(x, y) ? y : plb_curve_exact<CV>(x);
}

// Hot-path kernels: three channels, immediates, ONE (cold) fix-up branch for the group
template <int CV>
PLB_DEV void plb_curve3_s(float c[3])
{
    float y0, y1, y2;
    const bool h0 = plm_curve_s<CV>(c[0], y0), h1 = plm_curve_s<CV>(c[1], y1), h2 = plm_curve_s<CV>(c[2], y2);
    if (!(h0 & h1 & h2)) { // cold; static indices only (a dynamic one would put c[] in local memory)
        if (!h0) y0 = plb_curve_exact<CV>(c[0]);
        if (!h1) y1 = plb_curve_exact<CV>(c[1]);
        if (!h2) y2 = plb_curve_exact<CV>(c[2]);
    }
    c[0] = y0; c[1] = y1; c[2] = y2;
}

template <int CV>
PLB_DEV float plb_curve1_s(float x)
{
    float y;
    return plm_curve_s<CV>(x, y) ? y : plb_curve_exact<CV>(x);
}

// CV = PLB_CV_PQ_EOTF (plain) or PLB_CV_PQ_EOTF_SDR (result * from_pq, the "203 nits = 1" scale)
template <int CV>
PLB_DEV float plb_pq_eotf_cv(float x)
{
    return plb_curve<CV>(x);
}

// CV = PLB_CV_PQ_OETF (plain), _SDR (argument * to_pq) or _DELIN (argument * 1/from_pq)
template <int CV>
PLB_DEV float plb_pq_oetf_cv(float x, bool clamp0)
{
    return plb_curve<CV>(clamp0 ? fmaxf(x, 0.0f) : x);
}

PLB_DEV float plb_pq_eotf(float x, bool clamp0)
{
    return plb_pq_eotf_cv<PLB_CV_PQ_EOTF>(clamp0 ? fmaxf(x, 0.0f) : x);
}

PLB_DEV float plb_pq_oetf(float x, bool clamp0)
{
    return plb_pq_oetf_cv<PLB_CV_PQ_OETF>(x, clamp0);
}

// `trc` is passed beside the constants so that a caller that knows it at compile time (the
// specialised HDR output pass, k_pixel_hdr) gets the one case it needs and nothing else.
PLB_DEV void plb_linearize(const int trc, const plb_trc_consts &t, float c[3])
{
    if (trc == TRC_LINEAR)
        return;
    if (trc != TRC_SCRGB) {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = fmaxf(c[k], 0.0f);
    }
    bool post_scale = false;
    switch (trc) {
    case TRC_SRGB:
        for (int k = 0; k < 3; k++)
            c[k] = 0.04045f < c[k] ? plm_pow_cv<PLB_CV_POW_24>((c[k] + 0.055f) / 1.055f)
                                   : c[k] * (float) (1.0 / 12.92);
        post_scale = true;
        break;
    case TRC_BT_1886: {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = t.a * plm_pow_cv<PLB_CV_POW_24>(c[k] + t.b);
        break;
    }
    case TRC_GAMMA18: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], 1.8f); post_scale = true; break;
    case TRC_GAMMA20: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], 2.0f); post_scale = true; break;
    case TRC_UNKNOWN:
    case TRC_GAMMA22: for (int k = 0; k < 3; k++) c[k] = plm_pow_cv<PLB_CV_POW_22>(c[k]); post_scale = true; break;
    case TRC_GAMMA24: for (int k = 0; k < 3; k++) c[k] = plm_pow_cv<PLB_CV_POW_24>(c[k]); post_scale = true; break;
    case TRC_GAMMA26: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], 2.6f); post_scale = true; break;
    case TRC_GAMMA28: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], 2.8f); post_scale = true; break;
    case TRC_PRO_PHOTO:
        for (int k = 0; k < 3; k++)
            c[k] = 0.03125f < c[k] ? plm_powf(c[k], 1.8f) : c[k] * (float) (1.0 / 16.0);
        post_scale = true;
        break;
    case TRC_ST428:
        for (int k = 0; k < 3; k++)
            c[k] = (float) (52.37 / 48.0) * plm_powf(c[k], 2.6f);
        post_scale = true;
        break;
    case TRC_PQ:
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = plb_pq_eotf_cv<PLB_CV_PQ_EOTF_SDR>(c[k]);
        break;
    case TRC_HLG: {
        // t.a = y, t.b = b
        for (int k = 0; k < 3; k++) {
            float x = (1 - t.b) * c[k] + t.b;
            x = 0.5f < x ? plm_expf((x - c_pq.hlg_c) * c_pq.hlg_inv_a) + c_pq.hlg_b : 4.0f * x * x;
            c[k] = x * (float) (1.0 / 12.0);
        }
        float y = plb_dot3_fma(t.luma[0], t.luma[1], t.luma[2], c[0], c[1], c[2]);
        float g = t.csp_max * plm_powf(fmaxf(y, 0.0f), t.a - 1);
        for (int k = 0; k < 3; k++)
            c[k] *= g;
        break;
    }
    case TRC_SCRGB:
        for (int k = 0; k < 3; k++)
            c[k] *= t.a; // 80/203 literal
        break;
    case TRC_V_LOG:
        for (int k = 0; k < 3; k++)
            c[k] = 0.181f <= c[k] ? plm_powf(10.0f, (c[k] - c_pq.vlog_d) * c_pq.vlog_inv_c) - c_pq.vlog_b
                                  : (c[k] - 0.125f) * (float) (1.0 / 5.6);
        break;
    case TRC_S_LOG1:
        for (int k = 0; k < 3; k++)
            c[k] = plm_powf(10.0f, (c[k] - c_pq.slog_c) * c_pq.slog_inv_a) - c_pq.slog_b;
        break;
    case TRC_S_LOG2:
        for (int k = 0; k < 3; k++)
            c[k] = c_pq.slog_q <= c[k]
                 ? (plm_powf(10.0f, (c[k] - c_pq.slog_c) * c_pq.slog_inv_a) - c_pq.slog_b) * c_pq.slog_inv_k2
                 : (c[k] - c_pq.slog_q) * c_pq.slog_inv_p;
        break;
    default:
        break;
    }
    if (post_scale && (t.csp_max != 1 || t.csp_min != 0)) {
        for (int k = 0; k < 3; k++)
            c[k] = (t.csp_max - t.csp_min) * c[k] + t.csp_min;
    }
}

PLB_DEV void plb_delinearize(const int trc, const plb_trc_consts &t, float c[3])
{
    if (trc == TRC_LINEAR)
        return;
    if (t.black_scaled) {
        const float ps_a = 1 / (t.csp_max - t.csp_min), ps_b = -t.csp_min / (t.csp_max - t.csp_min);
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = ps_a * c[k] + ps_b;
    }
    if (trc != TRC_SCRGB) {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = fmaxf(c[k], 0.0f);
    }
    switch (trc) {
    case TRC_SRGB:
        for (int k = 0; k < 3; k++)
            c[k] = 0.0031308f <= c[k] ? 1.055f * plm_pow_cv<PLB_CV_POW_INV24>(c[k]) - 0.055f
                                      : c[k] * 12.92f;
        break;
    case TRC_BT_1886: { // t.a = 1/a, t.b = b
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = plm_pow_cv<PLB_CV_POW_INV24>(t.a * c[k]) - t.b;
        break;
    }
    case TRC_GAMMA18: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], (float) (1.0 / 1.8)); break;
    case TRC_GAMMA20: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], (float) (1.0 / 2.0)); break;
    case TRC_UNKNOWN:
    case TRC_GAMMA22: for (int k = 0; k < 3; k++) c[k] = plm_pow_cv<PLB_CV_POW_INV22>(c[k]); break;
    case TRC_GAMMA24: for (int k = 0; k < 3; k++) c[k] = plm_pow_cv<PLB_CV_POW_INV24>(c[k]); break;
    case TRC_GAMMA26: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], (float) (1.0 / 2.6)); break;
    case TRC_GAMMA28: for (int k = 0; k < 3; k++) c[k] = plm_powf(c[k], (float) (1.0 / 2.8)); break;
    case TRC_ST428:
        for (int k = 0; k < 3; k++)
            c[k] = plm_powf(c[k] * (float) (48.0 / 52.37), (float) (1.0 / 2.6));
        break;
    case TRC_PRO_PHOTO:
        for (int k = 0; k < 3; k++)
            c[k] = 0.001953f <= c[k] ? plm_powf(c[k], (float) (1.0 / 1.8)) : c[k] * 16.0f;
        break;
    case TRC_PQ:
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = plb_pq_oetf_cv<PLB_CV_PQ_OETF_DELIN>(c[k], false);
        break;
    case TRC_HLG: {
        for (int k = 0; k < 3; k++)
            c[k] *= 1.0f / t.csp_max;
        float y = plb_dot3_fma(t.luma[0], t.luma[1], t.luma[2], c[0], c[1], c[2]);
        float g = 12.0f * plm_powf(fmaxf(1e-6f, y), (1 - t.a) / t.a);
        for (int k = 0; k < 3; k++) {
            float x = c[k] * g;
            x = 1.0f < x ? c_pq.hlg_a * plm_logf(x - c_pq.hlg_b) + c_pq.hlg_c : 0.5f * sqrtf(x);
            c[k] = (1 / (1 - t.b)) * x + (-t.b / (1 - t.b));
        }
        break;
    }
    case TRC_SCRGB:
        for (int k = 0; k < 3; k++)
            c[k] *= t.a; // 203/80 literal
        break;
    case TRC_V_LOG:
        for (int k = 0; k < 3; k++)
            c[k] = 0.01f <= c[k] ? c_pq.vlog_c10 * plm_logf(c[k] + c_pq.vlog_b) + c_pq.vlog_d
                                 : 5.6f * c[k] + 0.125f;
        break;
    case TRC_S_LOG1:
        for (int k = 0; k < 3; k++)
            c[k] = c_pq.slog_a10 * plm_logf(c[k] + c_pq.slog_b) + c_pq.slog_c;
        break;
    case TRC_S_LOG2:
        for (int k = 0; k < 3; k++)
            c[k] = 0.0f <= c[k] ? c_pq.slog_a10 * plm_logf(c_pq.slog_k2 * c[k] + c_pq.slog_b) + c_pq.slog_c
                                : c_pq.slog_p * c[k] + c_pq.slog_q;
        break;
    default:
        break;
    }
}

PLB_DEV void plb_decode_special(int sys, float c[3])
{
    const float lms2rgb[9] = { 3.43661f, -0.79133f, -0.0259499f,
                               -2.50645f, 1.98360f, -0.0989137f,
                               0.06984f, -0.192271f, 1.12486f };
    switch (sys) {
    case PL_COLOR_SYSTEM_BT_2020_C: {
        float g = c[1];
        c[2] = c[2] * (c[2] <= 0.0f ? 1.9404f : 1.5816f) + g;
        c[0] = c[0] * (c[0] <= 0.0f ? 1.7184f : 0.9936f) + g;
        float lin[3];
        for (int k = 0; k < 3; k++)
            lin[k] = 0.08145f <= c[k]
                   ? plm_powf((c[k] + 0.0993f) * (float) (1.0 / 1.0993), (float) (1.0 / 0.45))
                   : c[k] * (float) (1.0 / 4.5);
        float gg = (lin[1] - 0.2627f * lin[0] - 0.0593f * lin[2]) * (float) (1.0 / 0.6780);
        c[1] = 0.0181f <= gg ? 1.0993f * plm_powf(gg, 0.45f) - 0.0993f : gg * 4.5f;
        break;
    }
    case PL_COLOR_SYSTEM_BT_2100_PQ:
        for (int k = 0; k < 3; k++)
            c[k] = plb_pq_eotf(c[k], true);
        plb_mat3_lit(lms2rgb, c[0], c[1], c[2]);
        for (int k = 0; k < 3; k++)
            c[k] = plb_pq_oetf(c[k], true);
        break;
    case PL_COLOR_SYSTEM_BT_2100_HLG:
        for (int k = 0; k < 3; k++)
            c[k] = 0.5f < c[k] ? plm_expf((c[k] - c_pq.hlg_c) * c_pq.hlg_inv_a) + c_pq.hlg_b
                               : 4.0f * c[k] * c[k];
        plb_mat3_lit(lms2rgb, c[0], c[1], c[2]);
        for (int k = 0; k < 3; k++)
            c[k] = 1.0f < c[k] ? c_pq.hlg_a * plm_logf(c[k] - c_pq.hlg_b) + c_pq.hlg_c
                               : 0.5f * sqrtf(c[k]);
        break;
    }
}

PLB_DEV void plb_encode_special(int sys, float c[3])
{
    const float rgb2lms[9] = { 0.412109f, 0.166748f, 0.024170f,
                               0.523925f, 0.720459f, 0.075440f,
                               0.063965f, 0.112793f, 0.900394f };
    switch (sys) {
    case PL_COLOR_SYSTEM_BT_2020_C: {
        float lin[3];
        for (int k = 0; k < 3; k++)
            lin[k] = 0.08145f <= c[k]
                   ? plm_powf((c[k] + 0.0993f) * (float) (1.0 / 1.0993), (float) (1.0 / 0.45))
                   : c[k] * (float) (1.0 / 4.5);
        float g = plb_dot3_fma(0.2627f, 0.6780f, 0.0593f, lin[0], lin[1], lin[2]);
        g = 0.0181f <= g ? 1.0993f * plm_powf(g, 0.45f) - 0.0993f : g * 4.5f;
        c[1] = g;
        c[2] = c[2] - g;
        c[0] = c[0] - g;
        c[2] *= c[2] <= 0.0f ? (float) (1.0 / 1.9404) : (float) (1.0 / 1.5816);
        c[0] *= c[0] <= 0.0f ? (float) (1.0 / 1.7184) : (float) (1.0 / 0.9936);
        break;
    }
    case PL_COLOR_SYSTEM_BT_2100_PQ:
        for (int k = 0; k < 3; k++)
            c[k] = plb_pq_eotf(c[k], true);
        plb_mat3_lit(rgb2lms, c[0], c[1], c[2]);
        for (int k = 0; k < 3; k++)
            c[k] = plb_pq_oetf(c[k], false);
        break;
    case PL_COLOR_SYSTEM_BT_2100_HLG:
        for (int k = 0; k < 3; k++)
            c[k] = 0.5f < c[k] ? plm_expf((c[k] - c_pq.hlg_c) * c_pq.hlg_inv_a) + c_pq.hlg_b
                               : 4.0f * c[k] * c[k];
        plb_mat3_lit(rgb2lms, c[0], c[1], c[2]);
        for (int k = 0; k < 3; k++)
            c[k] = 1.0f < c[k] ? c_pq.hlg_a * plm_logf(c[k] - c_pq.hlg_b) + c_pq.hlg_c
                               : 0.5f * sqrtf(c[k]);
        break;
    }
}

// Trilinear fetch from the gamut 3DLUT (rgba16 values pre-converted to float). Addressing as
// lut.c:285-295 (texel centres at (i + 0.5) / n, clamp to edge) in its reduced form: with the
// index saturated to [0, 1] the texel-space coordinate is x (n - 1); lerp(a, b, t) = fma(t, b - a, a).
// (The reference samples a hardware-filtered 3D texture: 8-bit interpolation weights.)
PLB_DEV float plb_lerp(float a, float b, float t) { return __fmaf_rn(t, b - a, a); }

PLB_DEV void plb_gamut_sample(const plb_op_color_map &p, const float idx[3], float out[3])
{
    const int n[3] = { p.lut_I, p.lut_C, p.lut_h };
    int i0[3], i1[3];
    float f[3];
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        const float t = __saturatef(idx[a]) * p.g_nm1[a]; // NaN -> 0
        const float fl = floorf(t);
        i0[a] = (int) fl;
        i1[a] = min(i0[a] + 1, n[a] - 1);
        f[a] = t - fl;
    }
    // the two texels along I are adjacent in memory; each (C, h) row pair is mixed along I as
    // soon as it arrives (same order as the GLSL trilinear filter: I, then C, then h)
    const float4 *lut = (const float4 *) p.gamut_lut;
    const int r00 = (i0[2] * n[1] + i0[1]) * n[0], r01 = (i0[2] * n[1] + i1[1]) * n[0];
    const int r10 = (i1[2] * n[1] + i0[1]) * n[0], r11 = (i1[2] * n[1] + i1[1]) * n[0];
    const float4 t000 = __ldg(lut + r00 + i0[0]), t001 = __ldg(lut + r00 + i1[0]);
    const float4 t010 = __ldg(lut + r01 + i0[0]), t011 = __ldg(lut + r01 + i1[0]);
    const float4 t100 = __ldg(lut + r10 + i0[0]), t101 = __ldg(lut + r10 + i1[0]);
    const float4 t110 = __ldg(lut + r11 + i0[0]), t111 = __ldg(lut + r11 + i1[0]);
#define PLB_TRI(F, k) { \
        const float x00 = plb_lerp(t000.F, t001.F, f[0]), x01 = plb_lerp(t010.F, t011.F, f[0]); \
        const float x10 = plb_lerp(t100.F, t101.F, f[0]), x11 = plb_lerp(t110.F, t111.F, f[0]); \
        const float y0 = plb_lerp(x00, x01, f[1]), y1 = plb_lerp(x10, x11, f[1]); \
        out[k] = plb_lerp(y0, y1, f[2]); }
    PLB_TRI(x, 0) PLB_TRI(y, 1) PLB_TRI(z, 2)
#undef PLB_TRI
}

// Bilinear fetch from the single-component fp16 feature map (textureLod on a linearly filtered
// texture, clamp to edge): the general texel path, or direct half loads when HOT.
template <bool HOT>
PLB_DEV float plb_fmap_linear(const plb_tex_desc &t, float u, float v)
{
    if (!HOT || t.kind != PLB_FK_HALF || t.comps != 1)
        return plb_tex_linear(t, u, v).x;
    const float tx = u * (float) t.w - 0.5f, ty = v * (float) t.h - 0.5f;
    const float bx = floorf(tx), by = floorf(ty);
    const float fx = tx - bx, fy = ty - by;
    const int x0 = plb_clampi((int) bx, 0, t.w - 1), x1 = plb_clampi((int) bx + 1, 0, t.w - 1);
    const int y0 = plb_clampi((int) by, 0, t.h - 1), y1 = plb_clampi((int) by + 1, 0, t.h - 1);
    const __half *r0 = (const __half *) ((const char *) t.ptr + (size_t) y0 * t.pitch);
    const __half *r1 = (const __half *) ((const char *) t.ptr + (size_t) y1 * t.pitch);
    const float a = __half2float(r0[x0]), b = __half2float(r0[x1]);
    const float c = __half2float(r1[x0]), d = __half2float(r1[x1]);
    return plb_mix(plb_mix(a, b, fx), plb_mix(c, d, fx), fy);
}

// pl_shader_custom_lut (shaders/lut.c:212-280). 1-D: every channel through its own component,
// linear in the table (lut.c:700-715 form); 3-D: tetrahedral interpolation (lut.c:773-809):
// the cube cell is cut into six tetrahedra along the black-white diagonal, the four vertices
// are fetched as texels and blended with barycentric weights in the order the GLSL adds them.
PLB_DEV float4 plb_lut3_texel(const float *lut, const int size[3], int x, int y, int z)
{
    return __ldg((const float4 *) lut + ((size_t) z * size[1] + y) * size[0] + x);
}

static __device__ __noinline__ void plb_custom_lut(const plb_color_op &op, float c[4])
{
    const float *lut = op.u.custom_lut.lut;
    const int *size = op.u.custom_lut.size;
    if (op.u.custom_lut.has_in)
        plb_matv(op.u.custom_lut.shaper_in, c[0], c[1], c[2]);
    if (!size[1]) {
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            const float fpos = plb_clamp(c[k], 0.0f, 1.0f) * (float) (size[0] - 1);
            const float fbase = floorf(fpos), fceil = ceilf(fpos);
            c[k] = plb_mix(lut[4 * (int) fbase + k], lut[4 * (int) fceil + k], fpos - fbase);
        }
    } else {
        float pos[3], base[3], f[3];
        int v0[3], v3[3], v1[3], v2[3];
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            pos[k] = plb_clamp(c[k], 0.0f, 1.0f) * (float) (size[k] - 1);
            base[k] = floorf(pos[k]);
            f[k] = pos[k] - base[k];
            v0[k] = v1[k] = (int) base[k];
            v3[k] = v2[k] = (int) ceilf(pos[k]);
        }
        // c_ab = fpart.a >= fpart.b for (x,y), (y,z), (z,x); the reversed tests are their negations
        const bool cxy = f[0] >= f[1], cyz = f[1] >= f[2], czx = f[2] >= f[0];
        const bool cyx = !cxy, czy = !cyz, cxz = !czx;
        float s[3] = { f[0], f[1], f[2] };
#define TETRA(X, Y, Z, CA, CB) \
        if ((CA) && (CB)) { s[0] = f[X]; s[1] = f[Y]; s[2] = f[Z]; v1[X] = v3[X]; v2[Z] = v0[Z]; }
        TETRA(0, 1, 2, cxy, cyz)
        TETRA(0, 2, 1, cxz, czy)
        TETRA(2, 0, 1, czx, cxy)
        TETRA(2, 1, 0, czy, cyx)
        TETRA(1, 2, 0, cyz, czx)
        TETRA(1, 0, 2, cyx, cxz)
#undef TETRA
        const float4 t0 = plb_lut3_texel(lut, size, v0[0], v0[1], v0[2]);
        const float4 t1 = plb_lut3_texel(lut, size, v1[0], v1[1], v1[2]);
        const float4 t2 = plb_lut3_texel(lut, size, v2[0], v2[1], v2[2]);
        const float4 t3 = plb_lut3_texel(lut, size, v3[0], v3[1], v3[2]);
        const float w0 = 1.0f - s[0], w1 = s[0] - s[1], w2 = s[1] - s[2], w3 = s[2];
        c[0] = w0 * t0.x + w1 * t1.x + w2 * t2.x + w3 * t3.x;
        c[1] = w0 * t0.y + w1 * t1.y + w2 * t2.y + w3 * t3.y;
        c[2] = w0 * t0.z + w1 * t1.z + w2 * t2.z + w3 * t3.z;
    }
    if (op.u.custom_lut.has_out)
        plb_matv(op.u.custom_lut.shaper_out, c[0], c[1], c[2]);
}

// `tone_map(x)` of pl_shader_color_map_ex: clip / linear fast paths or the 1-D LUT (:1826-1876)
PLB_DEV float plb_tone_map_x(const plb_op_color_map &p, float x)
{
    if (p.tone_kind == 1)
        return plb_clamp(x, p.in_min, p.in_max);
    if (p.tone_kind == 2) {
        x = p.lin_a * x + p.lin_b;
        x = plb_clamp(x, 0.0f, 1.0f);
        return (p.out_max - p.out_min) * x + p.out_min;
    }
    return plb_lut1d(p.tone_lut, p.tone_lut_size, p.tone_a * x + p.tone_b);
}

// HOT: static curve evaluation (specialised kernels); else descriptor-driven.
// (gx, gy): the invocation, for the feature-map position.
template <bool HOT = false>
PLB_DEV void plb_color_map(const plb_op_color_map &p, float px[3], int gx, int gy)
{
    const float lms2ipt[9] = { 0.4000f, 0.4000f, 0.2000f, 4.4550f, -4.8510f, 0.3960f,
                               0.8056f, 0.3572f, -1.1628f };
    const float ipt2lms[9] = { 1.0f, 0.0975689f, 0.205226f, 1.0f, -0.1138760f, 0.133217f,
                               1.0f, 0.0326151f, -0.676887f };
    float lms[3] = { px[0], px[1], px[2] };
    plb_matv(p.rgb2lms, lms[0], lms[1], lms[2]);
    float ipt[3];
    #pragma unroll
    for (int k = 0; k < 3; k++)
        ipt[k] = fmaxf(lms[k], 0.0f);
    if (HOT) {
        plb_curve3_s<PLB_CV_PQ_OETF_SDR>(ipt);
    } else {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            ipt[k] = plb_curve<PLB_CV_PQ_OETF_SDR>(ipt[k]);
    }
    plb_matv(lms2ipt, ipt[0], ipt[1], ipt[2]);
    const float i_orig = ipt[0];

    if (p.need_tone) {
        float x = ipt[0];
        if (p.has_fmap) {
            // HF detail from a bicubic interpolation of the low-pass map (colorspace.c:1886-1917)
            const float lpos[2] = { __fdiv_rn((float) gx + 0.5f, (float) p.out_w),
                                    __fdiv_rn((float) gy + 0.5f, (float) p.out_h) };
            float g[4], h[4];
            #pragma unroll
            for (int a = 0; a < 2; a++) {
                const float t = lpos[a] * p.fm_size[a] + 0.5f;
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
            const float ax = lpos[0] + p.fm_pt[0] * h[0], ay = lpos[1] + p.fm_pt[1] * h[1];
            const float bx = lpos[0] + p.fm_pt[0] * h[2], by = lpos[1] + p.fm_pt[1] * h[3];
            const float l00 = plb_fmap_linear<HOT>(p.fmap, ax, ay), l01 = plb_fmap_linear<HOT>(p.fmap, ax, by);
            const float l0 = l01 * (1.0f - g[1]) + l00 * g[1];
            const float l10 = plb_fmap_linear<HOT>(p.fmap, bx, ay), l11 = plb_fmap_linear<HOT>(p.fmap, bx, by);
            const float l1 = l11 * (1.0f - g[1]) + l10 * g[1];
            const float luma = l1 * (1.0f - g[0]) + l0 * g[0];
            const float highres = plb_clamp(ipt[0], 0.0f, 1.0f);
            const float lowres = plb_clamp(luma, 0.0f, 1.0f);
            const float detail = highres - lowres;
            const float base = plb_tone_map_x(p, highres);
            const float sharp = plb_tone_map_x(p, lowres) + detail;
            x = base * (1.0f - p.recovery) + sharp * p.recovery;
            x = plb_clamp(x, p.out_min, p.out_max);
        } else {
            x = plb_tone_map_x(p, x);
        }
        ipt[0] = x;
        float h0 = ((i_orig - 6.0f) * i_orig + 9.0f) * i_orig;
        float h1 = ((ipt[0] - 6.0f) * ipt[0] + 9.0f) * ipt[0];
        float s = fminf(i_orig / ipt[0], h1 / h0);
        ipt[1] *= s;
        ipt[2] *= s;
    }

    if (p.need_gamut) {
        float idx[3];
        idx[0] = p.g_a * ipt[0] + p.g_b;
        idx[1] = 2.0f * sqrtf(__fmaf_rn(ipt[1], ipt[1], ipt[2] * ipt[2]));
        idx[2] = c_pq.inv_2pi * plm_atan2_fast(ipt[2], ipt[1]) + 0.5f;
        plb_gamut_sample(p, idx, ipt);
        ipt[1] -= (float) (32768.0 / 65535.0);
        ipt[2] -= (float) (32768.0 / 65535.0);
    }

    plb_matv(ipt2lms, ipt[0], ipt[1], ipt[2]);
    #pragma unroll
    for (int k = 0; k < 3; k++)
        lms[k] = fmaxf(ipt[k], 0.0f);
    if (HOT) {
        plb_curve3_s<PLB_CV_PQ_EOTF_SDR>(lms);
    } else {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            lms[k] = plb_curve<PLB_CV_PQ_EOTF_SDR>(lms[k]);
    }
    plb_matv(p.lms2rgb, lms[0], lms[1], lms[2]);
    px[0] = lms[0]; px[1] = lms[1]; px[2] = lms[2];
}

PLB_DEV void plb_dither(const plb_op_dither &p, int x, int y, float c[4])
{
    if (p.fast_pow2) {
        // LUT of power-of-two size, no temporal rotation, plain rounding: (x + 0.5) / size,
        // fract and * size are exact in fp32, i.e. the LUT index is x mod size, y mod size
        const float bias = __ldg(p.lut + (y & (p.lut_size - 1)) * p.lut_size + (x & (p.lut_size - 1)));
        #pragma unroll
        for (int k = 0; k < 4; k++)
            c[k] = floorf(p.scale * c[k] + bias) * p.inv_scale;
        return;
    }
    int size = 0;
    if (p.method == PL_DITHER_BLUE_NOISE || p.method == PL_DITHER_ORDERED_LUT)
        size = p.lut_size;
    else if (p.method == PL_DITHER_ORDERED_FIXED)
        size = 16;
    const float scale = (float) ((1ull << p.new_depth) - 1);
    float bias = 0.0f, px = 0.0f, py = 0.0f;
    if (size) {
        const float inv_size = 1.0f / (float) size;
        px = ((float) x + 0.5f) * inv_size;
        py = ((float) y + 0.5f) * inv_size;
        px -= floorf(px);
        py -= floorf(py);
        if (p.temporal) {
            float nx = __fmaf_rn(p.rot[2], py, p.rot[0] * px) + 1.0f;
            float ny = __fmaf_rn(p.rot[3], py, p.rot[1] * px) + 1.0f;
            px = nx - floorf(nx);
            py = ny - floorf(ny);
        }
    }
    switch (p.method) {
    case PL_DITHER_WHITE_NOISE: {
        uint3 st = make_uint3((uint32_t) x, (uint32_t) y, p.temporal ? p.frame_index : 0);
        bias = plb_prng(st).x;
        break;
    }
    case PL_DITHER_ORDERED_FIXED: {
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
        bias = __ldg(p.lut + iy * size + ix);
        break;
    }
    }

    if (p.gamma != 1.0f && p.new_depth <= 4) {
        for (int k = 0; k < 4; k++) {
            float col = c[k];
            float col_lin = plm_powf(col, p.gamma);
            float low, high, offset;
            if (p.new_depth == 1) {
                low = 0.0f; high = 1.0f; offset = col_lin;
            } else {
                low = floorf(col * scale) / scale;
                high = ceilf(col * scale) / scale;
                float low_lin = plm_powf(low, p.gamma), high_lin = plm_powf(high, p.gamma);
                float range = high_lin - low_lin;
                offset = (col_lin - low_lin) / fmaxf(range, 1e-6f);
            }
            c[k] = offset > bias ? high : low;
        }
    } else {
        const float inv_scale = 1.0f / scale;
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            float v = scale * c[k] + bias;
            c[k] = floorf(v) * inv_scale;
        }
    }
}

PLB_DEV float plb_round_fbo(float v, int kind, int bits)
{
    if (kind == 0)
        return __half2float(__float2half_rn(v));
    if (kind == 1) {
        const float maxv = (float) ((1u << bits) - 1);
        return rintf(plb_clamp(v, 0.0f, 1.0f) * maxv) / maxv;
    }
    return v;
}

// PLB_OP_PEAK is handled by the kernel that owns the block-wide reduction; here we only
// compute the per-pixel 14-bit PQ luma (colorspace.c:1279-1288, exact float constants).
// HOT: PQ source, static curve evaluation (specialised kernels).
template <bool HOT = false>
PLB_DEV uint32_t plb_peak_ypq(const int trc, const plb_op_peak &p, const float c_in[4])
{
    float c[3] = { c_in[0], c_in[1], c_in[2] };
    if (HOT) {
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = fmaxf(c[k], 0.0f);
        plb_curve3_s<PLB_CV_PQ_EOTF_SDR>(c);
    } else {
        plb_linearize(trc, p.lin, c);
    }
    float luma = plb_dot3_fma(p.lin.luma[0], p.lin.luma[1], p.lin.luma[2], c[0], c[1], c[2]);
    luma *= 203.0f / 10000.0f;
    luma = plb_clamp(luma, 0.0f, 1.0f);
    luma = HOT ? plb_curve1_s<PLB_CV_PQX_OETF>(luma) : plb_curve<PLB_CV_PQX_OETF>(luma);
    if (p.cutoff != 0.0f) {
        float t = plb_clamp((luma - 0.0f) / (p.cutoff - 0.0f), 0.0f, 1.0f);
        luma *= t * t * (3.0f - 2.0f * t);
    }
    return (uint32_t) (16383.0f * luma);
}

// Runs all ops except PEAK (returned through peak_op for the caller to reduce).
// Deliberately NOT inlined: one copy per translation unit keeps every sampling kernel small
// (the call costs nothing next to the transcendental chains inside).
//
// The program is read from a per-CTA shared-memory copy (plb_s_prog, staged by
// plb_tail_stage): operands then come from LDS with link-time addresses instead of generic
// loads through a pointer to the kernel parameters.
static __shared__ plb_color_prog plb_s_prog;
static __shared__ plb_dovi_tab plb_s_dovi; // staged only when the program reshapes

// pl_shader_dovi_reshape, colorspace.c:106-273 (reshape_mmr :52-92, reshape_poly :94-97)
PLB_DEV float plb_dot4_fma(const float a[4], float x, float y, float z, float w)
{
    return __fmaf_rn(a[3], w, __fmaf_rn(a[2], z, __fmaf_rn(a[1], y, a[0] * x)));
}

PLB_DEV void plb_dovi_reshape(float c[4])
{
    const plb_dovi_tab &d = plb_s_dovi;
    const float sig[3] = { plb_clamp(c[0], 0.0f, 1.0f), plb_clamp(c[1], 0.0f, 1.0f),
                           plb_clamp(c[2], 0.0f, 1.0f) };
    #pragma unroll 1
    for (int ch = 0; ch < 3; ch++) {
        const int np = d.num_pivots[ch];
        if (!np)
            continue;
        float s = sig[ch];
        int k = 0;
        if (np > 2) {
            const float *pv = d.pivots[ch];
            // the mix tree of :206-216 (mix(a, b, bvec) takes b where true)
            if (s >= pv[3])
                k = s >= pv[5] ? (s >= pv[6] ? 7 : 6) : (s >= pv[4] ? 5 : 4);
            else
                k = s >= pv[1] ? (s >= pv[2] ? 3 : 2) : (s >= pv[0] ? 1 : 0);
        }
        const float *co = d.coeffs[ch][k];
        if (co[3] == 0.0f) {
            s = (co[2] * s + co[1]) * s + co[0];
        } else {
            const int idx = (int) co[1];
            const unsigned order = (unsigned) co[3];
            const float (*m)[4] = &d.mmr[ch][idx];
            const float x0 = sig[0] * sig[1], x1 = sig[0] * sig[2], x2 = sig[1] * sig[2];
            const float x3 = x0 * sig[2];
            s = co[0];
            s += plb_dot3_fma(m[0][0], m[0][1], m[0][2], sig[0], sig[1], sig[2]);
            s += plb_dot4_fma(m[1], x0, x1, x2, x3);
            if (order >= 2) {
                const float q0 = sig[0] * sig[0], q1 = sig[1] * sig[1], q2 = sig[2] * sig[2];
                const float y0 = x0 * x0, y1 = x1 * x1, y2 = x2 * x2, y3 = x3 * x3;
                s += plb_dot3_fma(m[2][0], m[2][1], m[2][2], q0, q1, q2);
                s += plb_dot4_fma(m[3], y0, y1, y2, y3);
                if (order >= 3) {
                    s += plb_dot3_fma(m[4][0], m[4][1], m[4][2], q0 * sig[0], q1 * sig[1], q2 * sig[2]);
                    s += plb_dot4_fma(m[5], y0 * x0, y1 * x1, y2 * x2, y3 * x3);
                }
            }
        }
        c[ch] = plb_clamp(s, d.lo[ch], d.hi[ch]);
    }
}

// "Light" programs contain only scale / matrix / alpha / swizzle / FBO-rounding steps (no
// transcendental): cheap enough to run inline, register-resident, in the resampling kernels
// (e.g. the output pass of an SDR upscale: 1/scale, swizzle). plb_prog_is_light() on the host
// decides; the semantics are those of the corresponding cases of plb_run_color_ops.
PLB_DEV void plb_light_step(const plb_color_op &op, const int kind, float c[4])
{
    if (kind == PLB_OP_SCALE) {
        const float s = op.u.scale.s;
        c[0] *= s; c[1] *= s; c[2] *= s;
        if (op.u.scale.with_alpha)
            c[3] *= s;
    } else if (kind == PLB_OP_MATRIX) {
        plb_matv(op.u.matrix.m, c[0], c[1], c[2]);
        if (op.u.matrix.has_c) {
            c[0] += op.u.matrix.c[0]; c[1] += op.u.matrix.c[1]; c[2] += op.u.matrix.c[2];
        }
    } else if (kind == PLB_OP_ROUND_FBO) {
        #pragma unroll
        for (int k = 0; k < 4; k++)
            c[k] = plb_round_fbo(c[k], op.u.round_fbo.kind, op.u.round_fbo.bits);
    } else if (kind == PLB_OP_ALPHA) {
        if (op.u.alpha.to == PL_ALPHA_NONE && op.u.alpha.from == PL_ALPHA_UNKNOWN) {
            c[3] = 1.0f;
        } else if (op.u.alpha.from == PL_ALPHA_INDEPENDENT && op.u.alpha.to == PL_ALPHA_PREMULTIPLIED) {
            c[0] *= c[3]; c[1] *= c[3]; c[2] *= c[3];
        } else if (op.u.alpha.from == PL_ALPHA_PREMULTIPLIED && op.u.alpha.to != PL_ALPHA_PREMULTIPLIED) {
            if (c[3] > 1e-6f) {
                c[0] /= c[3]; c[1] /= c[3]; c[2] /= c[3];
            }
            if (op.u.alpha.to == PL_ALPHA_NONE)
                c[3] = 1.0f;
        }
    } else if (kind == PLB_OP_SWIZZLE) {
        int sel[4]; // launch-uniform: source component of output k, -1 = default (0, 0, 0, 1)
        bool ident = true;
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const int m = op.u.swizzle.map[k];
            sel[k] = (k < op.u.swizzle.comps && m >= 0) ? m : -1;
            ident &= sel[k] < 0 || sel[k] == k;
        }
        const float o0 = c[0], o1 = c[1], o2 = c[2], o3 = c[3];
        if (ident) { // the usual RGB(A) target: nothing moves, absent components take defaults
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                if (sel[k] < 0)
                    c[k] = k == 3 ? 1.0f : 0.0f;
            }
        } else {
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const float dflt = k == 3 ? 1.0f : 0.0f;
                c[k] = sel[k] < 0 ? dflt : sel[k] == 0 ? o0 : sel[k] == 1 ? o1 : sel[k] == 2 ? o2 : o3;
            }
        }
        if (op.u.swizzle.force_alpha)
            c[3] = o3;
    }
}

PLB_DEV void plb_run_light_ops(float c[4])
{
    const plb_color_prog &prog = plb_s_prog;
    #pragma unroll 1
    for (int i = 0; i < prog.n; i++) {
        const plb_color_op &op = prog.ops[i];
        plb_light_step(op, op.kind, c);
    }
}

// The same program applied to N pixels in one pass over the op list: the kind dispatch and the
// swizzle predicates are launch-uniform and evaluated once per step instead of once per pixel;
// the per-pixel arithmetic of each step is exactly that of plb_run_light_ops. `prog` may live
// in the kernel parameters (constant bank) or in shared memory.
template <int N>
PLB_DEV void plb_run_light_ops_n(const plb_color_prog &prog, float c[N][4])
{
    #pragma unroll 1
    for (int i = 0; i < prog.n; i++) {
        const plb_color_op &op = prog.ops[i];
        const int kind = op.kind;
        if (kind == PLB_OP_SCALE) {
            const float s = op.u.scale.s;
            const bool wa = op.u.scale.with_alpha;
            #pragma unroll
            for (int p = 0; p < N; p++) {
                c[p][0] *= s; c[p][1] *= s; c[p][2] *= s;
                if (wa)
                    c[p][3] *= s;
            }
        } else if (kind == PLB_OP_MATRIX) {
            #pragma unroll
            for (int p = 0; p < N; p++) {
                plb_matv(op.u.matrix.m, c[p][0], c[p][1], c[p][2]);
                if (op.u.matrix.has_c) {
                    c[p][0] += op.u.matrix.c[0]; c[p][1] += op.u.matrix.c[1]; c[p][2] += op.u.matrix.c[2];
                }
            }
        } else if (kind == PLB_OP_ROUND_FBO) {
            #pragma unroll
            for (int p = 0; p < N; p++) {
                #pragma unroll
                for (int k = 0; k < 4; k++)
                    c[p][k] = plb_round_fbo(c[p][k], op.u.round_fbo.kind, op.u.round_fbo.bits);
            }
        } else if (kind == PLB_OP_ALPHA) {
            const int from = op.u.alpha.from, to = op.u.alpha.to;
            #pragma unroll
            for (int p = 0; p < N; p++) {
                if (to == PL_ALPHA_NONE && from == PL_ALPHA_UNKNOWN) {
                    c[p][3] = 1.0f;
                } else if (from == PL_ALPHA_INDEPENDENT && to == PL_ALPHA_PREMULTIPLIED) {
                    c[p][0] *= c[p][3]; c[p][1] *= c[p][3]; c[p][2] *= c[p][3];
                } else if (from == PL_ALPHA_PREMULTIPLIED && to != PL_ALPHA_PREMULTIPLIED) {
                    if (c[p][3] > 1e-6f) {
                        c[p][0] /= c[p][3]; c[p][1] /= c[p][3]; c[p][2] /= c[p][3];
                    }
                    if (to == PL_ALPHA_NONE)
                        c[p][3] = 1.0f;
                }
            }
        } else if (kind == PLB_OP_SWIZZLE) {
            int sel[4];
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int m = op.u.swizzle.map[k];
                sel[k] = (k < op.u.swizzle.comps && m >= 0) ? m : -1;
            }
            const bool fa = op.u.swizzle.force_alpha;
            bool ident = true;
            #pragma unroll
            for (int k = 0; k < 4; k++)
                ident &= sel[k] < 0 || sel[k] == k;
            if (ident) { // the usual RGB(A) target: nothing moves, absent components take defaults
                #pragma unroll
                for (int p = 0; p < N; p++) {
                    const float o3 = c[p][3];
                    #pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (sel[k] < 0)
                            c[p][k] = k == 3 ? 1.0f : 0.0f;
                    }
                    if (fa)
                        c[p][3] = o3;
                }
            } else {
                #pragma unroll
                for (int p = 0; p < N; p++) {
                    const float o0 = c[p][0], o1 = c[p][1], o2 = c[p][2], o3 = c[p][3];
                    #pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float dflt = k == 3 ? 1.0f : 0.0f;
                        c[p][k] = sel[k] < 0 ? dflt : sel[k] == 0 ? o0 : sel[k] == 1 ? o1 : sel[k] == 2 ? o2 : o3;
                    }
                    if (fa)
                        c[p][3] = o3;
                }
            }
        }
    }
}

static __device__ __noinline__ void plb_run_color_ops(int x, int y, float c[4],
                               const plb_op_peak **peak_op, uint32_t *y_pq)
{
    const plb_color_prog &prog = plb_s_prog;
    for (int i = 0; i < prog.n; i++) {
        const plb_color_op &op = prog.ops[i];
        switch (op.kind) {
        case PLB_OP_SCALE:
            c[0] *= op.u.scale.s; c[1] *= op.u.scale.s; c[2] *= op.u.scale.s;
            if (op.u.scale.with_alpha)
                c[3] *= op.u.scale.s;
            break;
        case PLB_OP_MATRIX: {
            plb_matv(op.u.matrix.m, c[0], c[1], c[2]);
            if (op.u.matrix.has_c) {
                c[0] += op.u.matrix.c[0]; c[1] += op.u.matrix.c[1]; c[2] += op.u.matrix.c[2];
            }
            break;
        }
        case PLB_OP_FEATURES: {
            // lms = (k * M) * rgb; PQ (clamped at 0); I = dot((0.4, 0.4, 0.2), lms)
            float lms[3] = { c[0], c[1], c[2] };
            plb_matv(op.u.features.m, lms[0], lms[1], lms[2]);
            #pragma unroll
            for (int k = 0; k < 3; k++)
                lms[k] = plb_pq_oetf(lms[k], true);
            c[0] = __fmaf_rn(0.2f, lms[2], __fmaf_rn(0.4f, lms[1], 0.4f * lms[0]));
            c[1] = 0.0f; c[2] = 0.0f; c[3] = 1.0f;
            break;
        }
        case PLB_OP_CUSTOM_LUT: plb_custom_lut(op, c); break;
        case PLB_OP_LINEARIZE:   plb_linearize(op.u.trc.trc, op.u.trc, c); break;
        case PLB_OP_DELINEARIZE: plb_delinearize(op.u.trc.trc, op.u.trc, c); break;
        case PLB_OP_DECODE_SPECIAL: plb_decode_special(op.u.special.sys, c); break;
        case PLB_OP_ENCODE_SPECIAL: plb_encode_special(op.u.special.sys, c); break;
        case PLB_OP_SIGMOID:
            for (int k = 0; k < 3; k++) {
                float v = plb_clamp(c[k], 0.0f, 1.0f);
                c[k] = op.u.sigmoid.center - (1.0f / op.u.sigmoid.slope) *
                       plm_logf(1.0f / (v * op.u.sigmoid.scale + op.u.sigmoid.offset) - 1.0f);
            }
            break;
        case PLB_OP_UNSIGMOID:
            for (int k = 0; k < 3; k++) {
                float v = plb_clamp(c[k], 0.0f, 1.0f);
                c[k] = (1.0f / op.u.sigmoid.scale) /
                       (1.0f + plm_expf(op.u.sigmoid.slope * (op.u.sigmoid.center - v))) -
                       op.u.sigmoid.offset / op.u.sigmoid.scale;
            }
            break;
        case PLB_OP_COLOR_MAP: plb_color_map<false>(op.u.cmap, c, x, y); break;
        case PLB_OP_DITHER:    plb_dither(op.u.dither, x, y, c); break;
        case PLB_OP_DOVI_RESHAPE: plb_dovi_reshape(c); break;
        case PLB_OP_DOVI_LMS:
            #pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = plb_pq_eotf(c[k], true);
            plb_matv(op.u.matrix.m, c[0], c[1], c[2]);
            #pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = plb_pq_oetf(c[k], true);
            break;
        case PLB_OP_ROUND_FBO:
            #pragma unroll
            for (int k = 0; k < 4; k++)
                c[k] = plb_round_fbo(c[k], op.u.round_fbo.kind, op.u.round_fbo.bits);
            break;
        case PLB_OP_ALPHA:
            // pl_shader_set_alpha, colorspace.c (independent <-> premultiplied)
            if (op.u.alpha.to == PL_ALPHA_NONE && op.u.alpha.from == PL_ALPHA_UNKNOWN) {
                c[3] = 1.0f;
            } else if (op.u.alpha.from == PL_ALPHA_INDEPENDENT && op.u.alpha.to == PL_ALPHA_PREMULTIPLIED) {
                c[0] *= c[3]; c[1] *= c[3]; c[2] *= c[3];
            } else if (op.u.alpha.from == PL_ALPHA_PREMULTIPLIED && op.u.alpha.to != PL_ALPHA_PREMULTIPLIED) {
                if (c[3] > 1e-6f) { // colorspace.c:34-38
                    c[0] /= c[3]; c[1] /= c[3]; c[2] /= c[3];
                }
                if (op.u.alpha.to == PL_ALPHA_NONE)
                    c[3] = 1.0f;
            }
            break;
        case PLB_OP_GAMMA:
            for (int k = 0; k < 3; k++)
                c[k] = plm_powf(fmaxf(c[k], 0.0f), op.u.gamma.g);
            break;
        case PLB_OP_SWIZZLE: {
            const float o[4] = { c[0], c[1], c[2], c[3] };
            c[0] = c[1] = c[2] = 0.0f;
            c[3] = 1.0f;
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k < op.u.swizzle.comps && op.u.swizzle.map[k] >= 0)
                    c[k] = o[op.u.swizzle.map[k]];
            }
            if (op.u.swizzle.force_alpha)
                c[3] = o[3];
            break;
        }
        case PLB_OP_PEAK:
            *peak_op = &op.u.peak;
            *y_pq = plb_peak_ypq<false>(op.u.peak.lin.trc, op.u.peak, c);
            break;
        default:
            break;
        }
    }
}
