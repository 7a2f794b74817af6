// Colour mapping and HDR peak detection on the CUDA backend: host side.
//   pl_shader_color_map_ex  src/shaders/colorspace.c:1612-2024 (host logic :1618-1790,
//                           LUT handling :1851-1876, :1939-1970)
//   pl_shader_detect_peak   colorspace.c:1155-1353
//   update_peak_buf / measure_peak  colorspace.c:1020-1153 (the only cross-frame state)
// The per-pixel part runs in csrc/color_ops.cuh (PLB_OP_COLOR_MAP) and pixel_core.cuh
// (PLB_OP_PEAK).
#define _GNU_SOURCE
#include "shader_priv.h"

#include <math.h>

void plb_record_linearize(pl_shader sh, const struct pl_color_space *csp);
void plb_record_delinearize(pl_shader sh, const struct pl_color_space *csp);
void plb_fill_trc(struct plb_trc_consts *t, const struct pl_color_space *csp, bool delin);

const struct pl_peak_detect_params pl_peak_detect_default_params = { PL_PEAK_DETECT_DEFAULTS };
const struct pl_peak_detect_params pl_peak_detect_high_quality_params = { PL_PEAK_DETECT_HQ_DEFAULTS };
const struct pl_color_map_params pl_color_map_default_params = { PL_COLOR_MAP_DEFAULTS };
const struct pl_color_map_params pl_color_map_high_quality_params = { PL_COLOR_MAP_HQ_DEFAULTS };

// ---------------------------------------------------------------------------------------
// peak detection state machine
// ---------------------------------------------------------------------------------------

enum {
    SLICES = 12,
    PQ_BITS = 14,
    PQ_MAX = (1 << PQ_BITS) - 1,
    HIST_BITS = 7,
    HIST_BIAS = 1 << (HIST_BITS - 1),
    HIST_BINS = (1 << HIST_BITS) - HIST_BIAS,
};
#define HIST_PQ(bin) (((bin) + HIST_BIAS) << (PQ_BITS - HIST_BITS))

struct peak_buf_data {
    unsigned frame_wg_count[SLICES];
    unsigned frame_wg_active[SLICES];
    unsigned frame_sum_pq[SLICES];
    unsigned frame_max_pq[SLICES];
    unsigned frame_hist[SLICES][HIST_BINS];
};

_Static_assert(sizeof(struct peak_buf_data) == PL_PEAK_BUF_SIZE, "peak buffer layout");

static bool peak_params_eq(const struct pl_peak_detect_params *a,
                           const struct pl_peak_detect_params *b)
{
    return a->smoothing_period == b->smoothing_period &&
           a->scene_threshold_low == b->scene_threshold_low &&
           a->scene_threshold_high == b->scene_threshold_high &&
           a->percentile == b->percentile;
}

// percentile of the merged 64-bin histogram, linear inside the bin
static float measure_peak(const struct peak_buf_data *data, float percentile)
{
    unsigned frame_max_pq = data->frame_max_pq[0];
    for (int k = 1; k < SLICES; k++)
        frame_max_pq = PLB_MAX(frame_max_pq, data->frame_max_pq[k]);
    const float frame_max = (float) frame_max_pq / PQ_MAX;
    if (percentile <= 0 || percentile >= 100)
        return frame_max;
    unsigned total = 0;
    for (int k = 0; k < SLICES; k++)
        for (int i = 0; i < HIST_BINS; i++)
            total += data->frame_hist[k][i];
    if (!total)
        return frame_max;
    const unsigned target = ceilf(percentile / 100.0f * total);
    if (target >= total)
        return frame_max;
    unsigned below = 0;
    for (int i = 0; i < HIST_BINS; i++) {
        unsigned upto = below;
        for (int k = 0; k < SLICES; k++)
            upto += data->frame_hist[k][i];
        if (upto < target) {
            below = upto;
            continue;
        }
        const unsigned count_low = below, count_high = upto + 1;
        const float pq_low = (float) HIST_PQ(i) / PQ_MAX;
        float pq_high = (float) HIST_PQ(i + 1) / PQ_MAX;
        if (count_high > total)
            pq_high = frame_max;
        const float ratio = (float) (target - count_low) / (count_high - count_low);
        return ratio * pq_high + (1 - ratio) * pq_low;
    }
    return frame_max;
}

static inline float smoothstepf(float e0, float e1, float x)
{
    if (e0 == e1)
        return x >= e0;
    x = (x - e0) / (e1 - e0);
    x = x < 0.0f ? 0.0f : x > 1.0f ? 1.0f : x;
    return x * x * (3.0f - 2.0f * x);
}

// Fold one frame's raw statistics into the smoothed (avg, max) pair
bool pl_peak_state_update(struct pl_peak_state *st, const void *peak_buf,
                          const struct pl_peak_detect_params *params)
{
    const struct peak_buf_data *data = peak_buf;
    params = params ? params : &pl_peak_detect_default_params;
    if (!(data->frame_wg_count[0] > 0))
        return false;
    uint64_t sum_pq = 0, wg_count = 0, wg_active = 0;
    for (int k = 0; k < SLICES; k++) {
        sum_pq += data->frame_sum_pq[k];
        wg_count += data->frame_wg_count[k];
        wg_active += data->frame_wg_active[k];
    }
    float avg_pq, max_pq;
    if (wg_active) {
        avg_pq = (float) sum_pq / (wg_active * PQ_MAX);
        max_pq = measure_peak(data, params->percentile);
    } else {
        avg_pq = max_pq = PL_COLOR_HDR_BLACK; // solid black frame
    }

    if (!st->avg_pq) {
        st->avg_pq = avg_pq;
        st->max_pq = max_pq;
    } else {
        // deviations below the measurement resolution are noise
        const float epsilon = 1.0f / PQ_MAX;
        if (fabsf(avg_pq - st->avg_pq) < epsilon)
            avg_pq = st->avg_pq;
        if (fabsf(max_pq - st->max_pq) < epsilon)
            max_pq = st->max_pq;
    }

    // first-order IIR low-pass
    const float coeff = params->smoothing_period ? 1.0f - expf(-1.0f / params->smoothing_period)
                                                 : 1.0f;
    st->avg_pq += coeff * (avg_pq - st->avg_pq);
    st->max_pq += coeff * (max_pq - st->max_pq);

    // scene change: snap towards the measurement when the average jumps
    if (params->scene_threshold_low > 0 && params->scene_threshold_high > 0) {
        const float log10_pq = 1e-2f;
        const float lo = params->scene_threshold_low * log10_pq;
        const float hi = params->scene_threshold_high * log10_pq;
        const float bias = (float) wg_active / wg_count;
        const float delta = bias * fabsf(avg_pq - st->avg_pq);
        const float mix = smoothstepf(lo, hi, delta);
        st->avg_pq = mix * avg_pq + (1 - mix) * st->avg_pq;
        st->max_pq = mix * max_pq + (1 - mix) * st->max_pq;
    }
    return true;
}

// read back a pending measurement (blocking: the reference's pl_buf_read is too)
static void update_peak_buf(pl_gpu gpu, struct plb_color_map_obj *obj)
{
    struct plb_peak_state *pk = &obj->peak;
    if (!pk->pending)
        return;
    struct peak_buf_data data;
    pk->pending = false;
    if (!plb_dev_download(gpu, &data, pk->dev_buf, sizeof(data))) {
        PLB_ERR(gpu->log, "Failed reading peak detection buffer!");
        return;
    }
    struct pl_peak_state st = { pk->avg_pq, pk->max_pq };
    if (!pl_peak_state_update(&st, &data, &pk->params)) {
        PLB_WARN(gpu->log, "Peak detection buffer was never written (shader not dispatched "
                 "before its result was requested?)");
        return;
    }
    pk->avg_pq = st.avg_pq;
    pk->max_pq = st.max_pq;
}

bool pl_shader_detect_peak(pl_shader sh, struct pl_color_space csp, pl_shader_obj *state,
                           const struct pl_peak_detect_params *params)
{
    params = params ? params : &pl_peak_detect_default_params;
    if (!plb_shader_require_color(sh, "pl_shader_detect_peak"))
        return false;
    pl_gpu gpu = sh->params.gpu;
    if (!plb_shader_try_compute(sh, 16, 16, true)) {
        PLB_ERR(sh->log, "HDR peak detection requires compute shaders!");
        return false;
    }
    struct plb_color_map_obj *obj = plb_shader_obj(sh, state, PLB_OBJ_COLOR_MAP);
    if (!obj)
        return false;
    struct plb_peak_state *pk = &obj->peak;
    if (peak_params_eq(&pk->params, params)) {
        update_peak_buf(gpu, obj); // consume the previous frame before overwriting
    } else {
        pl_reset_detected_peak(*state);
    }
    if (!pk->dev_buf) {
        pk->dev_buf = plb_dev_alloc(gpu, sizeof(struct peak_buf_data));
        if (!pk->dev_buf) {
            plb_shader_fail(sh, "Failed creating peak detection SSBO!");
            return false;
        }
    }
    if (!plb_dev_memset(gpu, pk->dev_buf, 0, sizeof(struct peak_buf_data)))
        return false;
    pk->params = *params;
    pk->pending = true;

    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_PEAK);
    if (!op)
        return false;
    pl_color_space_infer(&csp);
    plb_fill_trc(&op->u.peak.lin, &csp, false);
    op->u.peak.cutoff = fmaxf(params->black_cutoff, 0.0f) * 1e-2f;
    op->u.peak.use_histogram = params->percentile > 0 && params->percentile < 100;
    op->u.peak.buf = pk->dev_buf;
    return true;
}

bool pl_get_detected_hdr_metadata(const pl_shader_obj state, struct pl_hdr_metadata *out)
{
    if (!state || state->type != PLB_OBJ_COLOR_MAP)
        return false;
    struct plb_color_map_obj *obj = &state->u.cmap;
    update_peak_buf(state->gpu, obj);
    if (!obj->peak.avg_pq)
        return false;
    out->max_pq_y = obj->peak.max_pq;
    out->avg_pq_y = obj->peak.avg_pq;
    return true;
}

void pl_reset_detected_peak(pl_shader_obj state)
{
    if (!state || state->type != PLB_OBJ_COLOR_MAP)
        return;
    struct plb_peak_state *pk = &state->u.cmap.peak;
    uint32_t *buf = pk->dev_buf;
    memset(pk, 0, sizeof(*pk));
    pk->dev_buf = buf;
}

// ---------------------------------------------------------------------------------------
// colour mapping
// ---------------------------------------------------------------------------------------

static uint64_t fnv(uint64_t h, const void *data, size_t n)
{
    const uint8_t *p = data;
    for (size_t i = 0; i < n; i++)
        h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}

static uint64_t gamut_signature(const struct pl_gamut_map_params *g)
{
    uint64_t h = 0xcbf29ce484222325ull;
    h = fnv(h, g->function->name, strlen(g->function->name));
    h = fnv(h, &g->input_gamut, sizeof(g->input_gamut));
    h = fnv(h, &g->output_gamut, sizeof(g->output_gamut));
    h = fnv(h, &g->min_luma, sizeof(float));
    h = fnv(h, &g->max_luma, sizeof(float));
    h = fnv(h, &g->constants, sizeof(g->constants));
    h = fnv(h, &g->lut_size_I, 3 * sizeof(int));
    return h | 1;
}

static void put_mat(float out[9], const pl_matrix3x3 *m)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            out[3*i + j] = m->m[i][j];
}

void pl_shader_color_map_ex(pl_shader sh, const struct pl_color_map_params *params,
                            const struct pl_color_map_args *args)
{
    if (!plb_shader_require_color(sh, "pl_shader_color_map"))
        return;
    pl_gpu gpu = sh->params.gpu;
    struct pl_color_space src = args->src, dst = args->dst;
    struct plb_color_map_obj *obj = NULL;
    if (args->state) {
        pl_get_detected_hdr_metadata(*args->state, &src.hdr);
        obj = plb_shader_obj(sh, args->state, PLB_OBJ_COLOR_MAP);
        if (!obj)
            return;
    }

    pl_color_space_infer_map(&src, &dst);
    if (pl_color_space_equal(&src, &dst)) {
        if (args->prelinearized)
            plb_record_delinearize(sh, &dst);
        return;
    }

    params = params ? params : &pl_color_map_default_params;
    if (params->visualize_lut || params->show_clipping) {
        plb_shader_fail(sh, "pl_shader_color_map: debug visualisations are outside the CUDA "
                        "hot path");
        return;
    }
    if (params->lut3d_tricubic) {
        plb_shader_fail(sh, "pl_shader_color_map: tricubic 3DLUT interpolation is not "
                        "implemented on the CUDA backend");
        return;
    }

    struct pl_tone_map_params tone = {
        .function = PLB_DEF(params->tone_mapping_function, &pl_tone_map_clip),
        .constants = params->tone_constants,
        .input_scaling = PL_HDR_PQ,
        .output_scaling = PL_HDR_PQ,
        .lut_size = PLB_DEF(params->lut_size, pl_color_map_default_params.lut_size),
        .hdr = src.hdr,
    };
    pl_color_space_nominal_luma_ex(pl_nominal_luma_params(
        .color = &src,
        .metadata = params->metadata,
        .scaling = tone.input_scaling,
        .out_min = &tone.input_min,
        .out_max = &tone.input_max,
        .out_avg = &tone.input_avg,
    ));
    pl_color_space_nominal_luma_ex(pl_nominal_luma_params(
        .color = &dst,
        .metadata = PL_HDR_METADATA_HDR10,
        .scaling = tone.output_scaling,
        .out_min = &tone.output_min,
        .out_max = &tone.output_max,
    ));
    pl_tone_map_params_infer(&tone);

    // values that only differ by rounding are the same value
    if (fabs(tone.input_max - tone.output_max) < 1e-6)
        tone.output_max = tone.input_max;
    if (fabs(tone.input_min - tone.output_min) < 1e-6)
        tone.output_min = tone.input_min;
    if (!params->inverse_tone_mapping)
        tone.output_max = PLB_MIN(tone.output_max, tone.input_max);

    const int *lut3d_def = pl_color_map_default_params.lut3d_size;
    struct pl_gamut_map_params gamut = {
        .function = PLB_DEF(params->gamut_mapping, &pl_gamut_map_clip),
        .constants = params->gamut_constants,
        .input_gamut = src.hdr.prim,
        .output_gamut = dst.hdr.prim,
        .lut_size_I = PLB_DEF(params->lut3d_size[0], lut3d_def[0]),
        .lut_size_C = PLB_DEF(params->lut3d_size[1], lut3d_def[1]),
        .lut_size_h = PLB_DEF(params->lut3d_size[2], lut3d_def[2]),
        .lut_stride = 3,
    };
    pl_color_space_nominal_luma_ex(pl_nominal_luma_params(
        .color = &dst,
        .metadata = PL_HDR_METADATA_HDR10,
        .scaling = PL_HDR_PQ,
        .out_min = &gamut.min_luma,
        .out_max = &gamut.max_luma,
    ));
    // without gamut expansion, never leave the source gamut
    if (!params->gamut_expansion && gamut.function->bidirectional) {
        if (pl_primaries_compatible(&gamut.input_gamut, &gamut.output_gamut))
            gamut.output_gamut = pl_primaries_clip(&gamut.output_gamut, &gamut.input_gamut);
    }

    bool can_fast = !params->force_tone_mapping_lut;
    if (!args->state) {
        // no state object: only the closed-form paths are available
        can_fast = true;
        if (tone.function != &pl_tone_map_clip)
            tone.function = &pl_tone_map_linear;
        if (gamut.function != &pl_gamut_map_clip)
            gamut.function = &pl_gamut_map_saturation;
    }

    bool need_tone_map = !pl_tone_map_params_noop(&tone);
    bool need_gamut_map = !pl_gamut_map_params_noop(&gamut);

    if (!args->prelinearized)
        plb_record_linearize(sh, &src);

    pl_matrix3x3 rgb2lms = pl_ipt_rgb2lms(pl_raw_primaries_get(src.primaries));
    pl_matrix3x3 lms2rgb = pl_ipt_lms2rgb(pl_raw_primaries_get(dst.primaries));

    if (need_gamut_map && gamut.function == &pl_gamut_map_saturation && can_fast) {
        // saturation mapping is a plain change of primaries
        const pl_matrix3x3 lms2src = pl_ipt_lms2rgb(&gamut.input_gamut);
        const pl_matrix3x3 dst2lms = pl_ipt_rgb2lms(&gamut.output_gamut);
        pl_matrix3x3_mul(&lms2rgb, &dst2lms);
        pl_matrix3x3_mul(&lms2rgb, &lms2src);
        need_gamut_map = false;
    }

    if (!need_tone_map && !need_gamut_map) {
        if (src.primaries != dst.primaries) {
            pl_matrix3x3_mul(&lms2rgb, &rgb2lms);
            struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_MATRIX);
            if (!op)
                return;
            put_mat(op->u.matrix.m, &lms2rgb);
        }
        plb_record_delinearize(sh, &dst);
        return;
    }


    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_COLOR_MAP);
    if (!op)
        return;
    struct plb_op_color_map *cm = &op->u.cmap;
    put_mat(cm->rgb2lms, &rgb2lms);
    put_mat(cm->lms2rgb, &lms2rgb);
    cm->need_tone = need_tone_map;
    cm->need_gamut = need_gamut_map;

    if (need_tone_map) {
        cm->in_min = tone.input_min;
        cm->in_max = tone.input_max;
        cm->out_min = tone.output_min;
        cm->out_max = tone.output_max;
        if (tone.function == &pl_tone_map_clip && can_fast) {
            cm->tone_kind = 1;
        } else if (tone.function == &pl_tone_map_linear && can_fast) {
            const float gain = tone.constants.exposure;
            const float scale = tone.input_max - tone.input_min;
            cm->tone_kind = 2;
            cm->lin_a = gain / scale;
            cm->lin_b = -gain / scale * tone.input_min;
        } else {
            cm->tone_kind = 0;
            if (!obj->have_tone || !pl_tone_map_params_equal(&tone, &obj->tone_params) ||
                obj->tone_lut_size != tone.lut_size)
            {
                float *tmp = malloc(tone.lut_size * sizeof(float));
                pl_tone_map_generate(tmp, &tone);
                if (obj->tone_lut_size != tone.lut_size) {
                    plb_dev_free(gpu, obj->dev_tone_lut);
                    obj->dev_tone_lut = plb_dev_alloc(gpu, tone.lut_size * sizeof(float));
                    obj->tone_lut_size = tone.lut_size;
                }
                bool ok = obj->dev_tone_lut &&
                          plb_dev_upload(gpu, obj->dev_tone_lut, tmp, tone.lut_size * sizeof(float));
                free(tmp);
                if (!ok) {
                    plb_shader_fail(sh, "Failed generating tone-mapping LUT!");
                    return;
                }
                obj->tone_params = tone;
                obj->have_tone = true;
            }
            const float lut_range = tone.input_max - tone.input_min;
            cm->tone_lut = obj->dev_tone_lut;
            cm->tone_lut_size = tone.lut_size;
            cm->tone_a = 1.0f / lut_range;
            cm->tone_b = -tone.input_min / lut_range;
        }
    }

    // contrast recovery (colorspace.c:1879-1921): only when the curve compresses
    if (need_tone_map && tone.input_max >= tone.output_max && params->contrast_recovery &&
        args->feature_map) {
        pl_tex fm = args->feature_map;
        if (!fm->params.sampleable) {
            plb_shader_fail(sh, "pl_shader_color_map: feature map is not sampleable");
            return;
        }
        cm->has_fmap = 1;
        cm->fmap = plb_tex_priv(fm)->d;
        cm->recovery = params->contrast_recovery;
        cm->fm_size[0] = (float) fm->params.w;
        cm->fm_size[1] = (float) fm->params.h;
        cm->fm_pt[0] = (float) (1.0 / fm->params.w);
        cm->fm_pt[1] = (float) (1.0 / fm->params.h);
        plb_shader_use_tex(sh, fm);
    }

    if (need_gamut_map) {
        const size_t texels = (size_t) gamut.lut_size_I * gamut.lut_size_C * gamut.lut_size_h;
        const uint64_t sig = gamut_signature(&gamut);
        if (!obj->have_gamut || obj->gamut_sig != sig || obj->gamut_texels != texels) {
            // fill_gamut_lut, colorspace.c:1589-1610: quantised to rgba16 with the chroma biased
            // by 32767. The reference's texture unit converts each fetched texel with n/65535;
            // here that (exact IEEE) conversion happens once at upload, so the kernel fetches
            // the very same float values as float4 and spares 24 divisions per pixel.
            float *tmp = malloc(texels * 3 * sizeof(float));
            float *q = malloc(texels * 4 * sizeof(float));
            pl_gamut_map_generate(tmp, &gamut);
            for (size_t i = 0; i < texels; i++) {
                const uint16_t q0 = roundf(tmp[3*i + 0] * UINT16_MAX);
                const uint16_t q1 = roundf(tmp[3*i + 1] * UINT16_MAX + (UINT16_MAX >> 1));
                const uint16_t q2 = roundf(tmp[3*i + 2] * UINT16_MAX + (UINT16_MAX >> 1));
                q[4*i + 0] = (float) q0 / 65535.0f;
                q[4*i + 1] = (float) q1 / 65535.0f;
                q[4*i + 2] = (float) q2 / 65535.0f;
                q[4*i + 3] = 0.0f;
            }
            if (obj->gamut_texels != texels) {
                plb_dev_free(gpu, obj->dev_gamut_lut);
                obj->dev_gamut_lut = plb_dev_alloc(gpu, texels * 4 * sizeof(float));
                obj->gamut_texels = texels;
            }
            bool ok = obj->dev_gamut_lut &&
                      plb_dev_upload(gpu, obj->dev_gamut_lut, q, texels * 4 * sizeof(float));
            free(tmp);
            free(q);
            if (!ok) {
                plb_shader_fail(sh, "Failed generating gamut-mapping LUT!");
                return;
            }
            obj->gamut_sig = sig;
            obj->have_gamut = true;
        }
        const float lut_range = gamut.max_luma - gamut.min_luma;
        cm->gamut_lut = obj->dev_gamut_lut;
        cm->lut_I = gamut.lut_size_I;
        cm->lut_C = gamut.lut_size_C;
        cm->lut_h = gamut.lut_size_h;
        cm->g_a = 1.0f / lut_range;
        cm->g_b = -gamut.min_luma / lut_range;
        const int lut_n[3] = { cm->lut_I, cm->lut_C, cm->lut_h };
        for (int a = 0; a < 3; a++)
            cm->g_nm1[a] = (float) (lut_n[a] - 1);
    }

    plb_record_delinearize(sh, &dst);
}

void pl_shader_color_map(pl_shader sh, const struct pl_color_map_params *params,
                         struct pl_color_space src, struct pl_color_space dst,
                         pl_shader_obj *state, bool prelinearized)
{
    pl_shader_color_map_ex(sh, params, pl_color_map_args(
        .src = src,
        .dst = dst,
        .prelinearized = prelinearized,
        .state = state,
    ));
}
