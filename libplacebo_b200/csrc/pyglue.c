// Python binding glue: flat-argument wrappers over the PUBLIC C API (include/libplacebo_b200.h)
// for ctypes callers — libplacebo_b200/stream.py, bench.py. Nothing here touches library
// internals: it is the code an application embedding the library would write (create a pl_gpu,
// upload planes, fill two pl_frames, call pl_render_image, download the target).
//
// A plg_stream renders planar YUV frames (e.g. YUV420P10) to one packed RGBA target with a
// fixed configuration. Frames live in `ring` resident texture sets so that consecutive frames
// touch different HBM (bench.py sizes the ring beyond the 126 MB L2).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "libplacebo_b200.h"

#define API __attribute__((visibility("default")))
#define MAX_RING 16
#define MAX_PASSES 32

// Per-pass execution times. Samples (cudaEvent intervals) reach the info callback a few
// dispatches after the run they belong to, in order; run i of a pass produces sample i.
struct pass_stat {
    uint64_t signature;
    char desc[256];
    uint64_t runs;            // dispatches seen so far
    uint64_t seen;            // samples consumed so far
    uint64_t first_run;       // first run index inside the collection window
    uint64_t last_run;        // one past the last run inside the window (0 = still open)
    uint64_t sum_ns, count;   // over samples whose run index lies in the window
};

struct plg_stream {
    pl_log log;
    pl_gpu gpu;
    pl_renderer rr;
    int ring;
    int src_w, src_h, cw, ch, dst_w, dst_h;
    pl_tex planes[MAX_RING][3];
    pl_tex target[MAX_RING];
    struct pl_frame image, target_frame;
    struct pl_render_params params;
    struct pl_deband_params deband;
    struct pl_peak_detect_params peak;
    struct pl_dither_params dither;
    struct pl_color_map_params cmap;
    struct pass_stat passes[MAX_PASSES];
    int num_passes;
    bool collect;
};

static void log_cb(void *priv, enum pl_log_level level, const char *msg)
{
    (void) priv;
    if (level <= PL_LOG_WARN)
        fprintf(stderr, "[libplacebo_b200] %s\n", msg);
}

static void info_cb(void *priv, const struct pl_render_info *info)
{
    struct plg_stream *s = priv;
    const struct pl_dispatch_info *p = info->pass;
    struct pass_stat *st = NULL;
    for (int i = 0; i < s->num_passes; i++) {
        if (s->passes[i].signature == p->signature)
            st = &s->passes[i];
    }
    if (!st) {
        if (s->num_passes == MAX_PASSES)
            return;
        st = &s->passes[s->num_passes++];
        memset(st, 0, sizeof(*st));
        st->signature = p->signature;
        st->last_run = UINT64_MAX;
        st->first_run = s->collect ? 0 : UINT64_MAX;
        snprintf(st->desc, sizeof(st->desc), "%s", p->shader->description);
    }
    if (s->collect && st->first_run == UINT64_MAX)
        st->first_run = st->runs;
    st->runs++;
    // consume the samples that arrived since the last callback. The history holds the last
    // 256 samples; bookkeeping is exact while a pass has run fewer times than that (bench.py
    // keeps warmup + steps below it) and stops counting afterwards.
    const uint64_t n = (uint64_t) p->num_samples;
    const uint64_t fresh = st->seen < 256 && n > st->seen ? n - st->seen : 0;
    for (uint64_t j = 0; j < fresh; j++) {
        const uint64_t idx = st->seen + j;
        if (idx >= st->first_run && idx < st->last_run) {
            st->sum_ns += p->samples[idx];
            st->count++;
        }
    }
    st->seen += fresh;
}

API struct plg_stream *plg_create(int device, int literal_coords)
{
    struct plg_stream *s = calloc(1, sizeof(*s));
    s->log = pl_log_create(PL_API_VER, pl_log_params( .log_cb = log_cb, .log_level = PL_LOG_WARN ));
    s->gpu = pl_cuda_create(s->log, pl_cuda_params( .device = device, .literal_coords = literal_coords ));
    if (!s->gpu) {
        pl_log_destroy(&s->log);
        free(s);
        return NULL;
    }
    s->rr = pl_renderer_create(s->log, s->gpu);
    return s;
}

API void plg_destroy(struct plg_stream *s)
{
    if (!s)
        return;
    pl_renderer_destroy(&s->rr);
    for (int r = 0; r < MAX_RING; r++) {
        for (int i = 0; i < 3; i++)
            pl_tex_destroy(s->gpu, &s->planes[r][i]);
        pl_tex_destroy(s->gpu, &s->target[r]);
    }
    pl_cuda_destroy(&s->gpu);
    pl_log_destroy(&s->log);
    free(s);
}

API void *plg_cuda_stream(struct plg_stream *s) { return pl_cuda_stream(s->gpu); }
API uint64_t plg_launch_count(struct plg_stream *s) { return pl_cuda_launch_count(s->gpu); }
API void plg_finish(struct plg_stream *s) { pl_gpu_finish(s->gpu); }

static struct pl_color_space csp_from_flat(const float *f)
{
    // layout shared with the tests: prim, trc, hdr.prim[8], min_luma, max_luma, max_cll,
    // max_fall, scene_max[3], scene_avg, max_pq_y, avg_pq_y
    struct pl_color_space c = {
        .primaries = (int) f[0],
        .transfer = (int) f[1],
        .hdr = {
            .prim = { .red = {f[2], f[3]}, .green = {f[4], f[5]}, .blue = {f[6], f[7]},
                      .white = {f[8], f[9]} },
            .min_luma = f[10], .max_luma = f[11], .max_cll = f[12], .max_fall = f[13],
            .scene_max = { f[14], f[15], f[16] }, .scene_avg = f[17],
            .max_pq_y = f[18], .avg_pq_y = f[19],
        },
    };
    return c;
}

enum {
    O_SRC_W, O_SRC_H, O_CHROMA_W, O_CHROMA_H, O_DST_W, O_DST_H,
    O_SRC_SYS, O_SRC_LEVELS, O_SRC_SAMPLE_DEPTH, O_SRC_COLOR_DEPTH, O_SRC_BIT_SHIFT,
    O_DST_SYS, O_DST_LEVELS, O_DST_SAMPLE_DEPTH, O_DST_COLOR_DEPTH,
    O_CHROMA_LOC, O_DEBAND, O_PEAK, O_ALLOW_DELAYED, O_DITHER, O_CMAP_HQ, O_RING,
    O_TONE_FN, O_GAMUT_FN, O_GRAIN, O_DOVI, O_ERRDIFF, O_CONTRAST, O_COUNT
};

// AV1 grain parameters of libplacebo's own test fixture (src/tests/utils.h:240-262): medium
// luma grain with a 3-tap-lag AR filter, chroma grain scaled from luma.
static struct pl_av1_grain_data demo_av1_grain(bool overlap)
{
    return (struct pl_av1_grain_data) {
        .num_points_y = 6,
        .points_y = {{0, 4}, {27, 33}, {54, 55}, {67, 61}, {108, 71}, {255, 72}},
        .num_points_uv = {2, 2},
        .points_uv = {{{0, 64}, {255, 64}}, {{0, 64}, {255, 64}}},
        .scaling_shift = 11,
        .ar_coeff_lag = 3,
        .ar_coeffs_y = {4, 1, 3, 0, 1, -3, 8, -3, 7, -23, 1, -25,
                        0, -10, 6, -17, -4, 53, 36, 5, -5, -17, 8, 66},
        .ar_coeffs_uv = {
            {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 127},
            {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 127},
        },
        .ar_coeff_shift = 7,
        .uv_mult_luma = {64, 64},
        .overlap = overlap,
    };
}

// Returns 0 on success. `iopts` follows the enum above; formats by name ("r16", "rgba8", ...)
// Dolby Vision RPU data with the shape of a profile-5 stream: 8-piece polynomial luma reshaping,
// order-3 MMR chroma (the values are those of libplacebo's own test fixture, src/tests/utils.h:
// 279-323), with non-trivial YCC -> RGB / RGB -> LMS matrices and offsets.
static const struct pl_dovi_metadata *demo_dovi(void)
{
    static struct pl_dovi_metadata m;
    static bool init;
    if (!init) {
        m = (struct pl_dovi_metadata) {
            .nonlinear_offset = { 0.0625f, 0.5f, 0.5f },
            .nonlinear = {{{1.0f, 0.0f, 1.5f}, {1.0f, -0.17f, -0.58f}, {1.0f, 1.88f, 0.0f}}},
            .linear = {{{0.95f, 0.04f, 0.01f}, {0.02f, 0.96f, 0.02f}, {0.0f, 0.03f, 0.97f}}},
            .comp = {
                {
                    .num_pivots = 9,
                    .pivots = {0.0615835786, 0.129032254, 0.353861183, 0.604105592, 0.854349971,
                               0.890518069, 0.906158328, 0.913978517, 0.92082113},
                    .poly_coeffs = {
                        {-0.0488376617, 1.99335372, -2.41716385}, {-0.0141925812, 1.61829138, -1.53397191},
                        {0.157061458, 0.63640213, -0.11302495}, {0.25272119, 0.246226311, 0.27281332},
                        {0.951621532, -1.35507894, 1.18898678}, {6.41251612, -13.6188488, 8.07336903},
                        {13.467535, -29.1869125, 16.6612244}, {28.2321472, -61.8516273, 34.7264938},
                    },
                }, {
                    .num_pivots = 2, .pivots = {0.0, 1.0}, .method = {1}, .mmr_order = {3},
                    .mmr_constant = {-0.500733018},
                    .mmr_coeffs = {{
                        {1.08411026, 3.80807829, 0.0881733894, -3.23097038, -0.409078479, -1.31310081, 2.71297002},
                        {-0.241833091, -3.57880807, -0.108109117, 3.13198471, 0.869203091, 1.96561158, -9.30871677},
                        {-0.177356839, 1.48970401, 0.0908923149, -0.510447979, -0.687603354, -0.934977889, 12.3544884},
                    }},
                }, {
                    .num_pivots = 2, .pivots = {0.0, 1.0}, .method = {1}, .mmr_order = {3},
                    .mmr_constant = {-1.23833287},
                    .mmr_coeffs = {{
                        {3.52909589, 0.383154511, 5.50820637, -1.02094889, -6.36386824, 0.194121242, 0.64683497},
                        {-2.57899785, -0.626081586, -6.05729723, 2.29143763, 9.14653015, -0.0507702827, -4.17724133},
                        {0.705404401, 0.341412306, 2.98387456, -1.71712542, -4.91501331, 0.1465137, 6.38665438},
                    }},
                },
            },
        };
        init = true;
    }
    return &m;
}

API int plg_configure(struct plg_stream *s, const int *o, const float *src_csp,
                      const float *dst_csp, const char *plane_fmt, const char *dst_fmt,
                      const char *upscaler, const char *plane_upscaler)
{
    s->ring = o[O_RING] < 1 ? 1 : o[O_RING] > MAX_RING ? MAX_RING : o[O_RING];
    s->src_w = o[O_SRC_W]; s->src_h = o[O_SRC_H];
    s->cw = o[O_CHROMA_W]; s->ch = o[O_CHROMA_H];
    s->dst_w = o[O_DST_W]; s->dst_h = o[O_DST_H];
    pl_fmt pf = NULL, df = NULL;
    for (int i = 0; i < s->gpu->num_formats; i++) {
        if (!strcmp(s->gpu->formats[i]->name, plane_fmt)) pf = s->gpu->formats[i];
        if (!strcmp(s->gpu->formats[i]->name, dst_fmt)) df = s->gpu->formats[i];
    }
    if (!pf || !df)
        return -2;
    for (int r = 0; r < s->ring; r++) {
        for (int i = 0; i < 3; i++) {
            if (!pl_tex_recreate(s->gpu, &s->planes[r][i], pl_tex_params(
                    .w = i ? s->cw : s->src_w, .h = i ? s->ch : s->src_h, .format = pf,
                    .sampleable = true, .host_writable = true )))
                return -3;
        }
        if (!pl_tex_recreate(s->gpu, &s->target[r], pl_tex_params(
                .w = s->dst_w, .h = s->dst_h, .format = df, .renderable = true,
                .storable = true, .host_readable = true )))
            return -3;
    }

    s->image = (struct pl_frame) {
        .num_planes = 3,
        .repr = { .sys = o[O_SRC_SYS], .levels = o[O_SRC_LEVELS],
                  .bits = { o[O_SRC_SAMPLE_DEPTH], o[O_SRC_COLOR_DEPTH], o[O_SRC_BIT_SHIFT] } },
        .color = csp_from_flat(src_csp),
    };
    for (int i = 0; i < 3; i++) {
        s->image.planes[i] = (struct pl_plane) {
            .texture = s->planes[0][i], .components = 1, .component_mapping = { i } };
    }
    pl_frame_set_chroma_location(&s->image, o[O_CHROMA_LOC] ? o[O_CHROMA_LOC] : PL_CHROMA_LEFT);
    if (o[O_DOVI]) {
        s->image.repr.sys = PL_COLOR_SYSTEM_DOLBYVISION;
        s->image.repr.dovi = demo_dovi();
    }
    if (o[O_GRAIN]) {
        s->image.film_grain.type = PL_FILM_GRAIN_AV1;
        s->image.film_grain.params.av1 = demo_av1_grain(o[O_GRAIN] > 1);
    }
    s->target_frame = (struct pl_frame) {
        .num_planes = 1,
        .planes = {{ .texture = s->target[0], .components = df->num_components,
                     .component_mapping = { 0, 1, 2, 3 } }},
        .repr = { .sys = o[O_DST_SYS], .levels = o[O_DST_LEVELS],
                  .bits = { o[O_DST_SAMPLE_DEPTH], o[O_DST_COLOR_DEPTH], 0 } },
        .color = csp_from_flat(dst_csp),
    };

    s->deband = pl_deband_default_params;
    s->peak = o[O_PEAK] == 2 ? pl_peak_detect_high_quality_params : pl_peak_detect_default_params;
    s->peak.allow_delayed = o[O_ALLOW_DELAYED];
    s->dither = pl_dither_default_params;
    if (o[O_DITHER] > 1)
        s->dither.method = o[O_DITHER] - 1;
    s->cmap = o[O_CMAP_HQ] ? pl_color_map_high_quality_params : pl_color_map_default_params;
    if (!o[O_CONTRAST]) // 1: keep the preset's contrast recovery (HQ: 0.3 with the low-pass feature map)
        s->cmap.contrast_recovery = 0;
    if (o[O_TONE_FN])
        s->cmap.tone_mapping_function = pl_tone_map_functions[o[O_TONE_FN] - 1];
    if (o[O_GAMUT_FN])
        s->cmap.gamut_mapping = pl_gamut_map_functions[o[O_GAMUT_FN] - 1];
    s->params = (struct pl_render_params) {
        PL_RENDER_DEFAULTS
        .upscaler = upscaler && *upscaler ? pl_find_filter_config(upscaler, 0) : NULL,
        .plane_upscaler = plane_upscaler && *plane_upscaler ? pl_find_filter_config(plane_upscaler, 0) : NULL,
        .deband_params = o[O_DEBAND] ? &s->deband : NULL,
        .peak_detect_params = o[O_PEAK] ? &s->peak : NULL,
        .dither_params = o[O_DITHER] ? &s->dither : NULL,
        .color_map_params = &s->cmap,
        .error_diffusion = o[O_ERRDIFF] ? pl_error_diffusion_kernels[o[O_ERRDIFF] - 1] : NULL,
        .info_callback = info_cb,
        .info_priv = s,
    };
    return 0;
}

// Host -> device copy of one frame into ring slot `slot` (asynchronous on the gpu's stream;
// the host buffers must stay valid until plg_finish or a later synchronising call).
static void noop_cb(void *priv) { (void) priv; }

// Asynchronous uploads (a callback makes pl_tex_upload non-blocking, gpu.h:879-896): the host
// planes must stay valid until plg_finish (bench.py keeps a ring of pinned frames).
API int plg_upload(struct plg_stream *s, int slot, const void *y, const void *u, const void *v)
{
    const void *ptr[3] = { y, u, v };
    for (int i = 0; i < 3; i++) {
        if (!pl_tex_upload(s->gpu, pl_tex_transfer_params( .tex = s->planes[slot][i],
                                                           .ptr = (void *) ptr[i],
                                                           .callback = noop_cb )))
            return -1;
    }
    return 0;
}

API int plg_render(struct plg_stream *s, int slot, int out_slot, int frame_index)
{
    for (int i = 0; i < 3; i++)
        s->image.planes[i].texture = s->planes[slot][i];
    s->target_frame.planes[0].texture = s->target[out_slot];
    if (frame_index >= 0)
        pl_renderer_set_frame_index(s->rr, (uint8_t) frame_index);
    if (s->image.film_grain.type) // AV1: a new random seed with every frame
        s->image.film_grain.seed = (uint64_t) (frame_index >= 0 ? frame_index + 1 : 1);
    if (!pl_render_image(s->rr, &s->image, &s->target_frame, &s->params))
        return -1;
    return (int) pl_renderer_get_errors(s->rr).errors;
}

// Asynchronous download (a callback makes pl_tex_download non-blocking, gpu.h:879-910): `host`
// must stay valid until plg_finish / plg_join + stream synchronisation.
API int plg_download(struct plg_stream *s, int out_slot, void *host)
{
    return pl_tex_download(s->gpu, pl_tex_transfer_params( .tex = s->target[out_slot],
                                                           .ptr = host,
                                                           .callback = noop_cb )) ? 0 : -1;
}

API void plg_join(struct plg_stream *s) { pl_cuda_join(s->gpu); }

// ---- frame-parallel peak statistics (include/libplacebo_b200.h, "frame-parallel extensions")
API int plg_measure_peak(struct plg_stream *s, int slot, int frame_index,
                         void *out_buf /* PL_PEAK_BUF_SIZE */)
{
    for (int i = 0; i < 3; i++)
        s->image.planes[i].texture = s->planes[slot][i];
    if (frame_index >= 0)
        pl_renderer_set_frame_index(s->rr, (uint8_t) frame_index);
    if (s->image.film_grain.type)
        s->image.film_grain.seed = (uint64_t) (frame_index >= 0 ? frame_index + 1 : 1);
    return pl_renderer_measure_peak(s->rr, &s->image, &s->params, out_buf) ? 0 : -1;
}

// asynchronous variant: statistics are copied into `dev_buf` (device memory) on the gpu's stream
API int plg_measure_peak_dev(struct plg_stream *s, int slot, int frame_index, void *dev_buf)
{
    for (int i = 0; i < 3; i++)
        s->image.planes[i].texture = s->planes[slot][i];
    if (frame_index >= 0)
        pl_renderer_set_frame_index(s->rr, (uint8_t) frame_index);
    if (s->image.film_grain.type)
        s->image.film_grain.seed = (uint64_t) (frame_index >= 0 ? frame_index + 1 : 1);
    return pl_renderer_measure_peak_dev(s->rr, &s->image, &s->params, dev_buf) ? 0 : -1;
}

API int plg_small_copy(struct plg_stream *s, void *dst, const void *src, size_t size, void *stream)
{
    return pl_cuda_small_copy(s->gpu, dst, src, size, stream) ? 0 : -1;
}

API int plg_peak_update(struct plg_stream *s, float state[2], const void *buf, int *scene_change)
{
    struct pl_peak_state st = { .avg_pq = state[0], .max_pq = state[1] };
    bool used = pl_peak_state_update(&st, buf, &s->peak);
    (void) scene_change;
    state[0] = st.avg_pq;
    state[1] = st.max_pq;
    return used;
}

API void plg_set_peak(struct plg_stream *s, const float state[2])
{
    struct pl_peak_state st = { .avg_pq = state[0], .max_pq = state[1] };
    pl_renderer_set_peak_state(s->rr, state[0] > 0 ? &st : NULL);
}

API int plg_get_peak(struct plg_stream *s, float out[2])
{
    struct pl_hdr_metadata md = {0};
    if (!pl_renderer_get_hdr_metadata(s->rr, &md))
        return -1;
    out[0] = md.avg_pq_y;
    out[1] = md.max_pq_y;
    return 0;
}

API void plg_reset(struct plg_stream *s)
{
    pl_renderer_destroy(&s->rr);
    s->rr = pl_renderer_create(s->log, s->gpu);
}

// ---- per-pass bookkeeping ------------------------------------------------------------------
// Open (on = 1) or close (on = 0) the collection window. After closing, render at least one
// more frame (and plg_finish before it) so that the last in-window samples get delivered.
API void plg_collect(struct plg_stream *s, int on)
{
    for (int i = 0; i < s->num_passes; i++) {
        struct pass_stat *st = &s->passes[i];
        if (on) {
            st->first_run = st->runs;
            st->last_run = UINT64_MAX;
            st->sum_ns = st->count = 0;
        } else {
            st->last_run = st->runs;
        }
    }
    s->collect = on;
}

// Writes "count<TAB>mean_ns<TAB>description" lines for the collection window; returns the
// number of distinct passes seen
API int plg_pass_list(struct plg_stream *s, char *out, size_t n)
{
    size_t len = 0;
    out[0] = 0;
    for (int i = 0; i < s->num_passes; i++) {
        const struct pass_stat *st = &s->passes[i];
        int k = snprintf(out + len, n - len, "%llu\t%llu\t%s\n",
                         (unsigned long long) st->count,
                         (unsigned long long) (st->count ? st->sum_ns / st->count : 0), st->desc);
        if (k < 0 || (size_t) k >= n - len)
            break;
        len += k;
    }
    return s->num_passes;
}
