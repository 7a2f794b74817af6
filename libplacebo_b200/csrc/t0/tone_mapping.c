// Host-side tone-mapping curve generation (1-D LUTs sampled by the colour-map kernel).
// Behavioural restatement of the reference's src/tone_mapping.c: parameter inference
// (:91-128), LUT driver (:147-177), knee selection shared by the ST 2094 / spline curves
// (:231-270) and the curve families (:299-760). Arithmetic keeps the reference's precision
// and operand order; parity is pinned bit-exactly by tests/test_t0_tone_mapping.py.
#define _GNU_SOURCE
#include "../internal.h"

#include <math.h>

#define MIX(a, b, x) ((x) * (b) + (1 - (x)) * (a))

static inline float fclampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : x > hi ? hi : x; }

static inline float smoothstep(float e0, float e1, float x)
{
    if (e0 == e1)
        return x >= e0;
    x = (x - e0) / (e1 - e0);
    x = clampf(x, 0.0f, 1.0f);
    return x * x * (3.0f - 2.0f * x);
}

static void sanitize_constants(struct pl_tone_map_constants *c)
{
    const float eps = 1e-6f;
    c->knee_adaptation   = fclampf(c->knee_adaptation, 0.0f, 1.0f);
    c->knee_minimum      = fclampf(c->knee_minimum, eps, 0.5f - eps);
    c->knee_maximum      = fclampf(c->knee_maximum, 0.5f + eps, 1.0f - eps);
    c->knee_default      = fclampf(c->knee_default, c->knee_minimum, c->knee_maximum);
    c->knee_offset       = fclampf(c->knee_offset, 0.5f, 2.0f);
    c->slope_tuning      = fclampf(c->slope_tuning, 0.0f, 10.0f);
    c->slope_offset      = fclampf(c->slope_offset, 0.0f, 1.0f);
    c->spline_contrast   = fclampf(c->spline_contrast, 0.0f, 1.5f);
    c->reinhard_contrast = fclampf(c->reinhard_contrast, eps, 1.0f - eps);
    c->linear_knee       = fclampf(c->linear_knee, eps, 1.0f - eps);
    c->exposure          = fclampf(c->exposure, eps, 10.0f);
}

bool pl_tone_map_params_equal(const struct pl_tone_map_params *a, const struct pl_tone_map_params *b)
{
    return a->function == b->function && a->input_scaling == b->input_scaling &&
           a->output_scaling == b->output_scaling && a->lut_size == b->lut_size &&
           a->input_min == b->input_min && a->input_max == b->input_max &&
           a->input_avg == b->input_avg && a->output_min == b->output_min &&
           a->output_max == b->output_max &&
           !memcmp(&a->constants, &b->constants, sizeof(a->constants)) &&
           pl_hdr_metadata_equal(&a->hdr, &b->hdr);
}

bool pl_tone_map_params_noop(const struct pl_tone_map_params *p)
{
    const float in_min = pl_hdr_rescale(p->input_scaling, PL_HDR_NITS, p->input_min);
    const float in_max = pl_hdr_rescale(p->input_scaling, PL_HDR_NITS, p->input_max);
    const float out_min = pl_hdr_rescale(p->output_scaling, PL_HDR_NITS, p->output_min);
    const float out_max = pl_hdr_rescale(p->output_scaling, PL_HDR_NITS, p->output_max);
    const bool can_inverse = p->function->map_inverse;
    return fabs(in_min - out_min) < 1e-4 &&      // black points agree
           in_max < out_max + 1e-2 &&            // nothing to compress
           (out_max < in_max + 1e-2 || !can_inverse); // nothing to expand
}

void pl_tone_map_params_infer(struct pl_tone_map_params *par)
{
    if (!par->function)
        par->function = &pl_tone_map_clip;
    sanitize_constants(&par->constants);

    // never treat the source as darker than SDR reference (100 nits) or than the target
    float sdr = pl_hdr_rescale(par->output_scaling, par->input_scaling, par->output_max);
    sdr = fminf(sdr, pl_hdr_rescale(PL_HDR_NITS, par->input_scaling, 100));
    par->input_max = fmaxf(par->input_max, sdr);

    // one-way curves cannot brighten
    if (!par->function->map_inverse) {
        const float in_max = pl_hdr_rescale(par->input_scaling, par->output_scaling, par->input_max);
        par->output_max = fminf(par->output_max, in_max);
    }
}

// parameters converted into the curve's native scaling
static struct pl_tone_map_params native_params(const struct pl_tone_map_params *params)
{
    struct pl_tone_map_params p = *params;
    pl_tone_map_params_infer(&p);
    const enum pl_hdr_scaling s = p.function->scaling;
    p.input_scaling = p.output_scaling = s;
    p.input_min  = pl_hdr_rescale(params->input_scaling, s, p.input_min);
    p.input_max  = pl_hdr_rescale(params->input_scaling, s, p.input_max);
    p.input_avg  = pl_hdr_rescale(params->input_scaling, s, p.input_avg);
    p.output_min = pl_hdr_rescale(params->output_scaling, s, p.output_min);
    p.output_max = pl_hdr_rescale(params->output_scaling, s, p.output_max);
    return p;
}

static void apply_curve(float *lut, const struct pl_tone_map_params *p)
{
    if (p->output_max > p->input_max + 1e-4)
        p->function->map_inverse(lut, p);
    else
        p->function->map(lut, p);
}

void pl_tone_map_generate(float *out, const struct pl_tone_map_params *params)
{
    const struct pl_tone_map_params p = native_params(params);
    const enum pl_hdr_scaling s = p.function->scaling;
    for (size_t i = 0; i < params->lut_size; i++) {
        float x = (float) i / (params->lut_size - 1);
        x = MIX(params->input_min, params->input_max, x);
        out[i] = pl_hdr_rescale(params->input_scaling, s, x);
    }
    apply_curve(out, &p);
    for (size_t i = 0; i < params->lut_size; i++) {
        const float x = clampf(out[i], p.output_min, p.output_max);
        out[i] = pl_hdr_rescale(s, params->output_scaling, x);
    }
}

float pl_tone_map_sample(float x, const struct pl_tone_map_params *params)
{
    struct pl_tone_map_params p = native_params(params);
    const enum pl_hdr_scaling s = p.function->scaling;
    p.lut_size = 1;
    x = clampf(x, params->input_min, params->input_max);
    x = pl_hdr_rescale(params->input_scaling, s, x);
    apply_curve(&x, &p);
    x = clampf(x, p.output_min, p.output_max);
    return pl_hdr_rescale(s, params->output_scaling, x);
}

// ---- shared helpers ----------------------------------------------------------------------------

#define LUT_LOOP(x) for (float *it_ = lut, *end_ = lut + p->lut_size, x; \
                         it_ < end_ && (x = *it_, 1); *it_++ = x)

static inline float norm_in(float x, const struct pl_tone_map_params *p)
{
    return (x - p->input_min) / (p->input_max - p->input_min);
}

// relative to the OUTPUT range (peak > 1 means "brighter than the display")
static inline float norm_rel(float x, const struct pl_tone_map_params *p)
{
    return (x - p->input_min) / (p->output_max - p->output_min);
}

static inline float denorm_out(float x, const struct pl_tone_map_params *p)
{
    return x * (p->output_max - p->output_min) + p->output_min;
}

static inline float bt1886_eotf(float x, float lo, float hi)
{
    const float lb = powf(lo, 1/2.4f), lw = powf(hi, 1/2.4f);
    return powf((lw - lb) * x + lb, 2.4f);
}

static inline float bt1886_oetf(float x, float lo, float hi)
{
    const float lb = powf(lo, 1/2.4f), lw = powf(hi, 1/2.4f);
    return (powf(x, 1/2.4f) - lb) / (lw - lb);
}

// Where to bend: source knee = scene average (or a default fraction of the range), target knee
// = blend between "keep the source brightness" and "keep the relative position", in PQ
static void pick_knee(float *src_out, float *dst_out, const struct pl_tone_map_params *p)
{
    const float src_min = pl_hdr_rescale(p->input_scaling, PL_HDR_PQ, p->input_min);
    const float src_max = pl_hdr_rescale(p->input_scaling, PL_HDR_PQ, p->input_max);
    const float src_avg = pl_hdr_rescale(p->input_scaling, PL_HDR_PQ, p->input_avg);
    const float dst_min = pl_hdr_rescale(p->output_scaling, PL_HDR_PQ, p->output_min);
    const float dst_max = pl_hdr_rescale(p->output_scaling, PL_HDR_PQ, p->output_max);
    const float kmin = p->constants.knee_minimum, kmax = p->constants.knee_maximum,
                kdef = p->constants.knee_default;

    const float src_lo = MIX(src_min, src_max, kmin), src_hi = MIX(src_min, src_max, kmax);
    const float dst_lo = MIX(dst_min, dst_max, kmin), dst_hi = MIX(dst_min, dst_max, kmax);
    float src_knee = PLB_DEF(src_avg, MIX(src_min, src_max, kdef));
    src_knee = fclampf(src_knee, src_lo, src_hi);

    const float target = (src_knee - src_min) / (src_max - src_min);
    const float adapted = MIX(dst_min, dst_max, target);
    // adapt fully only for extreme knee positions
    const float tuning = 1.0f - smoothstep(kmax, kdef, target) * smoothstep(kmin, kdef, target);
    const float adaptation = MIX(p->constants.knee_adaptation, 1.0f, tuning);
    float dst_knee = MIX(src_knee, adapted, adaptation);
    dst_knee = fclampf(dst_knee, dst_lo, dst_hi);

    *src_out = pl_hdr_rescale(PL_HDR_PQ, p->input_scaling, src_knee);
    *dst_out = pl_hdr_rescale(PL_HDR_PQ, p->output_scaling, dst_knee);
}

// ---- curves --------------------------------------------------------------------------------------

static void identity(float *lut, const struct pl_tone_map_params *p) { (void) lut; (void) p; }

// binomial coefficients C(n, k), n <= 16
static float binomial(int n, int k)
{
    double c = 1;
    for (int i = 1; i <= k; i++)
        c = c * (n - k + i) / i;
    return (float) (uint16_t) (c + 0.5);
}

static float bezier_intercept(uint8_t N, float Kx, float Ky)
{
    if (Kx <= 0 || Ky >= 1)
        return 1.0f / N;
    const float slope = Ky / Kx * (1 - Kx) / (1 - Ky);
    return fminf(slope / N, 1.0f);
}

// SMPTE ST 2094-40 Annex B: linear segment up to the knee, Bernstein polynomial above
static void st2094_40(float *lut, const struct pl_tone_map_params *p)
{
    const float D = p->output_max;
    float P[17], Kx, Ky, T;
    uint8_t N;
    if (p->hdr.ootf.num_anchors) {
        Kx = clampf(p->hdr.ootf.knee_x, 0, 1);
        Ky = clampf(p->hdr.ootf.knee_y, 0, 1);
        T = clampf(p->hdr.ootf.target_luma, p->input_min, p->input_max);
        N = p->hdr.ootf.num_anchors + 1;
        memcpy(P + 1, p->hdr.ootf.anchors, (N - 1) * sizeof(*P));
        P[0] = 0.0f;
        P[N] = 1.0f;
    } else {
        float src_knee, dst_knee;
        pick_knee(&src_knee, &dst_knee, p);
        Kx = src_knee / p->input_max;
        Ky = dst_knee / p->output_max;
        const float slope = Ky / Kx * (1 - Kx) / (1 - Ky);
        const int n = (int) ceilf(slope);
        N = n < 2 ? 2 : n > 16 ? 16 : n;
        P[0] = 0.0f;
        P[1] = bezier_intercept(N, Kx, Ky);
        for (int i = 2; i <= N; i++)
            P[i] = 1.0f;
        T = D;
    }

    if (D < T) {
        // display darker than the reference target: pull the knee towards black
        const float Dmin = 0.0f, u = fmaxf(0.0f, (D - Dmin) / (T - Dmin));
        Kx *= u;
        Ky *= u;
        const float beta = N * Kx / (1 - Kx);
        const float Kxy = fminf(Kx * p->input_max / D, beta / (beta + 1));
        Ky = MIX(Kxy, Ky, u);
        for (int i = 2; i <= N; i++)
            P[i] = MIX(1.0f, P[i], u);
        P[1] = MIX(bezier_intercept(N, Kx, Ky), P[1], u);
    } else if (D > T) {
        // display brighter: fade towards the identity
        const float w = powf(1 - (D - T) / (p->input_max - T), 1.4f);
        Ky *= T / D;
        const float Kxy = Kx * D / p->input_max;
        Ky = MIX(Kxy, Ky, w);
        for (int i = 2; i < N; i++) {
            const float lin = (float) i / N;
            P[i] = MIX(lin, P[i], w);
        }
        P[1] = MIX(bezier_intercept(N, Kx, Ky), P[1], w);
    }

    float coef[17];
    for (int i = 0; i <= N; i++)
        coef[i] = binomial(N, i);

    LUT_LOOP(x) {
        x = bt1886_oetf(x, p->input_min, p->input_max);
        x = bt1886_eotf(x, 0.0f, 1.0f);
        if (x <= Kx && Kx) {
            x *= Ky / Kx;
        } else {
            const float t = (x - Kx) / (1 - Kx);
            x = 0;
            for (uint8_t i = 0; i <= N; i++)
                x += coef[i] * powf(t, i) * powf(1 - t, N - i) * P[i];
            x = Ky + (1 - Ky) * x;
        }
        x = bt1886_oetf(x, 0.0f, 1.0f);
        x = bt1886_eotf(x, p->output_min, p->output_max);
    }
}

// SMPTE ST 2094-10 Annex B.2: rational curve through (min, knee, max)
static void st2094_10(float *lut, const struct pl_tone_map_params *p)
{
    float src_knee, dst_knee;
    pick_knee(&src_knee, &dst_knee, p);
    const float x1 = p->input_min, x2 = src_knee, x3 = p->input_max;
    const float y1 = p->output_min, y2 = dst_knee, y3 = p->output_max;
    const pl_matrix3x3 cmat = {{
        { x2*x3*(y2 - y3), x1*x3*(y3 - y1), x1*x2*(y1 - y2) },
        { x3*y3 - x2*y2,   x1*y1 - x3*y3,   x2*y2 - x1*y1   },
        { x3 - x2,         x1 - x3,         x2 - x1         },
    }};
    float c[3] = { y1, y2, y3 };
    pl_matrix3x3_apply(&cmat, c);
    const float k = 1.0 / (x3*y3*(x1 - x2) + x2*y2*(x3 - x1) + x1*y1*(x2 - x3));
    const float c1 = k * c[0], c2 = k * c[1], c3 = k * c[2];
    LUT_LOOP(x)
        x = (c1 + c2 * x) / (1 + c3 * x);
}

// ITU-R BT.2390 EETF: hermite roll-off above the knee start, black lift below
static void bt2390(float *lut, const struct pl_tone_map_params *p)
{
    const float minLum = norm_in(p->output_min, p), maxLum = norm_in(p->output_max, p);
    const float offset = p->constants.knee_offset;
    const float ks = (1 + offset) * maxLum - offset;
    const float bp = minLum > 0 ? fminf(1 / minLum, 4) : 4;
    const float gain_inv = 1 + minLum / maxLum * powf(1 - maxLum, bp);
    const float gain = maxLum < 1 ? 1 / gain_inv : 1;
    LUT_LOOP(x) {
        x = norm_in(x, p);
        if (ks < 1) {
            const float tb = (x - ks) / (1 - ks), tb2 = tb * tb, tb3 = tb2 * tb;
            const float pb = (2 * tb3 - 3 * tb2 + 1) * ks + (tb3 - 2 * tb2 + tb) * (1 - ks) +
                             (-2 * tb3 + 3 * tb2) * maxLum;
            x = x < ks ? x : pb;
        }
        if (x < 1) {
            x += minLum * powf(1 - x, bp);
            x = gain * (x - minLum) + minLum;
        }
        x = x * (p->input_max - p->input_min) + p->input_min;
    }
}

// ITU-R BT.2446 method A and its published inverse
static void bt2446a(float *lut, const struct pl_tone_map_params *p)
{
    const float phdr = 1 + 32 * powf(p->input_max / 10000, 1/2.4f);
    const float psdr = 1 + 32 * powf(p->output_max / 10000, 1/2.4f);
    LUT_LOOP(x) {
        x = powf(norm_in(x, p), 1/2.4f);
        x = logf(1 + (phdr - 1) * x) / logf(phdr);
        if (x <= 0.7399f)
            x = 1.0770f * x;
        else if (x < 0.9909f)
            x = (-1.1510f * x + 2.7811f) * x - 0.6302f;
        else
            x = 0.5f * x + 0.5f;
        x = (powf(psdr, x) - 1) / (psdr - 1);
        x = bt1886_eotf(x, p->output_min, p->output_max);
    }
}

static void bt2446a_inv(float *lut, const struct pl_tone_map_params *p)
{
    LUT_LOOP(x) {
        x = bt1886_oetf(x, p->input_min, p->input_max);
        x *= 255.0;
        if (x > 70)
            x = powf(x, (2.8305e-6f * x - 7.4622e-4f) * x + 1.2528f);
        else
            x = powf(x, (1.8712e-5f * x - 2.7334e-3f) * x + 1.3141f);
        x = powf(x / 1000, 2.4f);
        x = denorm_out(x, p);
    }
}

// Two polynomial segments (quadratic below, cubic above) meeting at the knee with a slope
// chosen from the compression ratio
static void spline(float *lut, const struct pl_tone_map_params *p)
{
    float src_pivot, dst_pivot;
    pick_knee(&src_pivot, &dst_pivot, p);
    float slope = (dst_pivot - p->output_min) / (src_pivot - p->input_min);
    float ratio = p->input_max / p->output_max - 1.0f;
    ratio = fclampf(p->constants.slope_tuning * ratio, p->constants.slope_offset,
                    1.0f + p->constants.slope_offset);
    slope = powf(slope, (1.0f - p->constants.spline_contrast) * ratio);

    const float in_min = p->input_min - src_pivot, in_max = p->input_max - src_pivot;
    const float out_min = p->output_min - dst_pivot, out_max = p->output_max - dst_pivot;
    const float Pa = (out_min - slope * in_min) / (in_min * in_min), Pb = slope;
    const float t = 2 * in_max * in_max;
    const float Qa = (slope * in_max - out_max) / (in_max * t);
    const float Qb = -3 * (slope * in_max - out_max) / t;
    const float Qc = slope;
    LUT_LOOP(x) {
        x -= src_pivot;
        x = x > 0 ? ((Qa * x + Qb) * x + Qc) * x : (Pa * x + Pb) * x;
        x += dst_pivot;
    }
}

static void reinhard(float *lut, const struct pl_tone_map_params *p)
{
    const float peak = norm_rel(p->input_max, p), contrast = p->constants.reinhard_contrast,
                offset = (1.0 - contrast) / contrast, scale = (peak + offset) / peak;
    LUT_LOOP(x) {
        x = norm_rel(x, p);
        x = x / (x + offset);
        x *= scale;
        x = denorm_out(x, p);
    }
}

static void mobius(float *lut, const struct pl_tone_map_params *p)
{
    const float peak = norm_rel(p->input_max, p), j = p->constants.linear_knee;
    const float a = -j*j * (peak - 1.0f) / (j*j - 2.0f * j + peak);
    const float b = (j*j - 2.0f * j * peak + peak) / fmaxf(1e-6f, peak - 1.0f);
    const float scale = (b*b + 2.0f * b*j + j*j) / (b - a);
    LUT_LOOP(x) {
        x = norm_rel(x, p);
        x = x <= j ? x : scale * (x + a) / (x + b);
        x = denorm_out(x, p);
    }
}

static inline float hable_fn(float x)
{
    const float A = 0.15, B = 0.50, C = 0.10, D = 0.20, E = 0.02, F = 0.30;
    return ((x * (A*x + C*B) + D*E) / (x * (A*x + B) + D*F)) - E/F;
}

static void hable(float *lut, const struct pl_tone_map_params *p)
{
    const float peak = p->input_max / p->output_max, scale = 1.0f / hable_fn(peak);
    LUT_LOOP(x) {
        x = bt1886_oetf(x, p->input_min, p->input_max);
        x = bt1886_eotf(x, 0, peak);
        x = scale * hable_fn(x);
        x = bt1886_oetf(x, 0, 1);
        x = bt1886_eotf(x, p->output_min, p->output_max);
    }
}

static void gamma_knee(float *lut, const struct pl_tone_map_params *p)
{
    const float peak = norm_rel(p->input_max, p), cutoff = p->constants.linear_knee,
                gamma = logf(cutoff) / logf(cutoff / peak);
    LUT_LOOP(x) {
        x = norm_rel(x, p);
        x = x > cutoff ? powf(x / peak, gamma) : x;
        x = denorm_out(x, p);
    }
}

static void linear_stretch(float *lut, const struct pl_tone_map_params *p)
{
    const float gain = p->constants.exposure;
    LUT_LOOP(x) {
        x = norm_in(x, p);
        x *= gain;
        x = denorm_out(x, p);
    }
}

#define TMF(sym, nm, desc, sc, fwd, inv) \
    const struct pl_tone_map_function sym = { .name = nm, .description = desc, .scaling = sc, \
                                              .map = fwd, .map_inverse = inv }

TMF(pl_tone_map_clip, "clip", "No tone mapping (clip)", PL_HDR_NORM, identity, identity);
TMF(pl_tone_map_st2094_40, "st2094-40", "SMPTE ST 2094-40 Annex B", PL_HDR_NITS, st2094_40, NULL);
TMF(pl_tone_map_st2094_10, "st2094-10", "SMPTE ST 2094-10 Annex B.2", PL_HDR_NITS, st2094_10, NULL);
TMF(pl_tone_map_bt2390, "bt2390", "ITU-R BT.2390 EETF", PL_HDR_PQ, bt2390, NULL);
TMF(pl_tone_map_bt2446a, "bt2446a", "ITU-R BT.2446 Method A", PL_HDR_NITS, bt2446a, bt2446a_inv);
TMF(pl_tone_map_spline, "spline", "Single-pivot polynomial spline", PL_HDR_PQ, spline, spline);
TMF(pl_tone_map_reinhard, "reinhard", "Reinhard", PL_HDR_NORM, reinhard, NULL);
TMF(pl_tone_map_mobius, "mobius", "Mobius", PL_HDR_NORM, mobius, NULL);
TMF(pl_tone_map_hable, "hable", "Filmic tone-mapping (Hable)", PL_HDR_NORM, hable, NULL);
TMF(pl_tone_map_gamma, "gamma", "Gamma function with knee", PL_HDR_NORM, gamma_knee, NULL);
TMF(pl_tone_map_linear, "linear", "Perceptually linear stretch", PL_HDR_PQ, linear_stretch, linear_stretch);
TMF(pl_tone_map_linear_light, "linearlight", "Linear light stretch", PL_HDR_NORM, linear_stretch, linear_stretch);

const struct pl_tone_map_function * const pl_tone_map_functions[] = {
    &pl_tone_map_clip, &pl_tone_map_st2094_40, &pl_tone_map_st2094_10, &pl_tone_map_bt2390,
    &pl_tone_map_bt2446a, &pl_tone_map_spline, &pl_tone_map_reinhard, &pl_tone_map_mobius,
    &pl_tone_map_hable, &pl_tone_map_gamma, &pl_tone_map_linear, &pl_tone_map_linear_light,
    NULL
};
const int pl_num_tone_map_functions = PLB_ARRAY_SIZE(pl_tone_map_functions) - 1;

const struct pl_tone_map_function *pl_find_tone_map_function(const char *name)
{
    for (int i = 0; name && i < pl_num_tone_map_functions; i++) {
        if (strcmp(name, pl_tone_map_functions[i]->name) == 0)
            return pl_tone_map_functions[i];
    }
    return NULL;
}
