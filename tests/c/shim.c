/*
 * Test driver for libplacebo_b200: plain C written against the PUBLIC header only, the way
 * the reference's own tests (src/tests/gpu_tests.c, tests/dummy.c) drive libplacebo. pytest
 * loads it through ctypes and only passes flat arguments, so no struct layout is ever
 * duplicated in Python.
 *
 *   plt_*   flat wrappers over the host-side T0 API (same source as the reference-side
 *           hvt_* wrappers in oracle/harvest.c: oracle/t0_flat.inc)
 *   shim_*  GPU drivers: create gpu/textures, build a pl_shader with the public pl_shader_*
 *           calls, pl_dispatch_finish it, download
 */
#include <libplacebo_b200.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
#define PL_DEF(x, d) ((x) ? (x) : (d))

static pl_log shim_log;

static void log_cb(void *priv, enum pl_log_level level, const char *msg)
{
    (void) priv;
    fprintf(stderr, "[plb:%d] %s\n", (int) level, msg);
}

static void shim_init(void)
{
    if (!shim_log)
        shim_log = pl_log_create(PL_API_VER, pl_log_params(
            .log_cb = log_cb,
            .log_level = getenv("PLB_DEBUG") ? PL_LOG_DEBUG : PL_LOG_WARN,
        ));
}

#define T0_HOT_PATH_ONLY 1
#define T0_FN(n) plt_##n
#define T0_API API
#define T0_INIT() shim_init()
#define T0_LOG shim_log
#define T0_GPU pl_cuda_format_catalogue()
#include "../../oracle/t0_flat.inc"

/* ---- gpu / textures ---------------------------------------------------------------------- */

/* shader state objects persist across calls so that LUT caching paths are exercised */
struct shim_objs {
    pl_shader_obj sampler, cmap, dither, grain[2];
};

struct shim_ctx {
    pl_gpu gpu;
    pl_dispatch dp;
    pl_renderer rr;
    struct shim_objs objs;
    /* planar target of the next shim_render call (shim_set_target_planes) */
    int tgt_nplanes, tgt_chroma_loc;
    pl_tex tgt_planes[3];
    /* deinterlacing of the next shim_render call (shim_set_deinterlace) */
    int di_field, di_first, di_algo, di_skip;
    pl_tex di_prev[3], di_next[3];
    /* frame mix of the next shim_render call (shim_set_mix): further frames around the vsync */
    int mix_n;
    pl_tex mix_planes[8][3];
    float mix_ts[8], mix_vsync;
    uint64_t mix_sig[8];
    char mix_filter[48];
    /* custom LUTs of the next shim_render call (shim_set_lut): 0 params, 1 image, 2 target */
    struct pl_custom_lut lut[3];
    int lut_type[3];
    pl_shader_obj lut_obj;
};

API void *shim_create(int literal_coords)
{
    shim_init();
    pl_gpu gpu = pl_cuda_create(shim_log, pl_cuda_params( .literal_coords = literal_coords ));
    if (!gpu)
        return NULL;
    struct shim_ctx *c = calloc(1, sizeof(*c));
    c->gpu = gpu;
    c->dp = pl_dispatch_create(shim_log, gpu);
    return c;
}

API void shim_reset_objs(struct shim_ctx *c);

API void shim_destroy(struct shim_ctx *c)
{
    if (!c)
        return;
    shim_reset_objs(c);
    pl_shader_obj_destroy(&c->lut_obj);
    pl_renderer_destroy(&c->rr);
    pl_dispatch_destroy(&c->dp);
    pl_cuda_destroy(&c->gpu);
    free(c);
}

API void shim_set_frame_index(struct shim_ctx *c, int index)
{
    pl_dispatch_set_frame_index(c->dp, (uint8_t) index);
}

API unsigned long long shim_launch_count(struct shim_ctx *c)
{
    return pl_cuda_launch_count(c->gpu);
}

/* data = tightly packed texels or NULL */
API void *shim_tex_create(struct shim_ctx *c, const char *fmt, int w, int h, const void *data)
{
    pl_fmt f = pl_find_named_fmt(c->gpu, fmt);
    if (!f)
        return NULL;
    return (void *) pl_tex_create(c->gpu, pl_tex_params(
        .w = w,
        .h = h,
        .format = f,
        .sampleable = true,
        .renderable = true,
        .storable = true,
        .host_writable = true,
        .host_readable = true,
        .blit_src = true,
        .blit_dst = true,
        .initial_data = data,
    ));
}

/* pl_tex_blit: rc = x0, y0, x1, y1 (all zero = whole texture) */
API void shim_tex_blit(struct shim_ctx *c, void *src, void *dst, const int *src_rc, const int *dst_rc,
                       int linear)
{
    pl_tex_blit(c->gpu, pl_tex_blit_params(
        .src = src, .dst = dst,
        .src_rc = { .x0 = src_rc[0], .y0 = src_rc[1], .x1 = src_rc[2], .y1 = src_rc[3] },
        .dst_rc = { .x0 = dst_rc[0], .y0 = dst_rc[1], .x1 = dst_rc[2], .y1 = dst_rc[3] },
        .sample_mode = linear ? PL_TEX_SAMPLE_LINEAR : PL_TEX_SAMPLE_NEAREST,
    ));
}

/* pl_frame_clear_rgba on a 3-plane YCbCr frame (sys / levels / depth given) or a packed RGB one */
API void shim_frame_clear(struct shim_ctx *c, int nplanes, void **planes, int sys, int levels,
                          int depth, const float *csp, const float *rgba)
{
    struct pl_frame f = {
        .num_planes = nplanes,
        .repr = { .sys = sys, .levels = levels, .bits = { depth, depth, 0 } },
        .color = csp_from_flat(csp),
    };
    for (int i = 0; i < nplanes; i++) {
        f.planes[i].texture = planes[i];
        if (nplanes == 1) {
            f.planes[i].components = 3;
            f.planes[i].component_mapping[0] = 0;
            f.planes[i].component_mapping[1] = 1;
            f.planes[i].component_mapping[2] = 2;
        } else {
            f.planes[i].components = 1;
            f.planes[i].component_mapping[0] = i;
        }
    }
    pl_frame_clear_rgba(c->gpu, &f, rgba);
}

API void shim_tex_destroy(struct shim_ctx *c, void *tex)
{
    pl_tex t = tex;
    pl_tex_destroy(c->gpu, &t);
}

API int shim_tex_upload(struct shim_ctx *c, void *tex, const void *data)
{
    return pl_tex_upload(c->gpu, pl_tex_transfer_params( .tex = tex, .ptr = (void *) data )) ? 0 : -1;
}

API int shim_tex_download(struct shim_ctx *c, void *tex, void *out)
{
    return pl_tex_download(c->gpu, pl_tex_transfer_params( .tex = tex, .ptr = out )) ? 0 : -1;
}

API void shim_finish(struct shim_ctx *c) { pl_gpu_finish(c->gpu); }

/* ---- shader drivers ------------------------------------------------------------------------ */

/*
 * One generic driver: a source stage followed by colour ops, dispatched into `dst`.
 * src_kind: 0 direct, 1 nearest, 2 bilinear, 3 deband, 4 polar, 5 ortho (one pass)
 * sargs:    source arguments (see below)
 * ops:      list of op codes terminated by 0; each op consumes floats from `oargs`
 *   1 decode_color   sys, levels, sample_depth, color_depth, bit_shift
 *   2 encode_color   sys, levels, sample_depth, color_depth, bit_shift
 *   3 linearize      csp[20]
 *   4 delinearize    csp[20]
 *   5 color_map      src csp[20], dst csp[20], preset(0 default/1 hq), tone_fn idx (-1 keep),
 *                    gamut_fn idx (-1 keep), use_state
 *   6 dither         depth, method, lut_size, temporal, transfer
 *   7 sigmoidize / 8 unsigmoidize   center, slope
 *   9 detect_peak    csp[20], hq   (statistics land in out_peak if given)
 */
API void shim_reset_objs(struct shim_ctx *c)
{
    pl_shader_obj_destroy(&c->objs.sampler);
    pl_shader_obj_destroy(&c->objs.cmap);
    pl_shader_obj_destroy(&c->objs.dither);
    pl_shader_obj_destroy(&c->objs.grain[0]);
    pl_shader_obj_destroy(&c->objs.grain[1]);
}

/* Dolby Vision metadata of the reference's own test fixture (src/tests/utils.h:279-323).
 * variant 0: as in the fixture (identity matrices); variant 1: same curves with a non-trivial
 * `linear` / `nonlinear` matrix and offsets, to exercise the matrix paths as well. */
static const struct pl_dovi_metadata *shim_dovi_fixture(int variant)
{
    static struct pl_dovi_metadata m[2];
    static bool init;
    if (!init) {
        const struct pl_dovi_metadata base = {
            .nonlinear = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}},
            .linear    = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}},
            .comp = {
                {
                    .num_pivots = 9,
                    .pivots = {0.0615835786, 0.129032254, 0.353861183,
                               0.604105592, 0.854349971, 0.890518069,
                               0.906158328, 0.913978517, 0.92082113},
                    .method = {0, 0, 0, 0, 0, 0, 0, 0},
                    .poly_coeffs = {
                        {-0.0488376617, 1.99335372, -2.41716385},
                        {-0.0141925812, 1.61829138, -1.53397191},
                        { 0.157061458, 0.63640213, -0.11302495},
                        {0.25272119, 0.246226311, 0.27281332},
                        {0.951621532, -1.35507894, 1.18898678},
                        {6.41251612, -13.6188488, 8.07336903},
                        {13.467535, -29.1869125, 16.6612244},
                        {28.2321472, -61.8516273, 34.7264938}
                    },
                }, {
                    .num_pivots = 2,
                    .pivots = {0.0, 1.0},
                    .method = {1},
                    .mmr_order = {3},
                    .mmr_constant = {-0.500733018},
                    .mmr_coeffs = {{
                        {1.08411026, 3.80807829, 0.0881733894, -3.23097038, -0.409078479, -1.31310081, 2.71297002},
                        {-0.241833091, -3.57880807, -0.108109117, 3.13198471, 0.869203091, 1.96561158, -9.30871677},
                        {-0.177356839, 1.48970401, 0.0908923149, -0.510447979, -0.687603354, -0.934977889, 12.3544884},
                    }},
                }, {
                    .num_pivots = 2,
                    .pivots = {0.0, 1.0},
                    .method = {1},
                    .mmr_order = {3},
                    .mmr_constant = {-1.23833287},
                    .mmr_coeffs = {{
                        {3.52909589, 0.383154511, 5.50820637, -1.02094889, -6.36386824, 0.194121242, 0.64683497},
                        {-2.57899785, -0.626081586, -6.05729723, 2.29143763, 9.14653015, -0.0507702827, -4.17724133},
                        {0.705404401, 0.341412306, 2.98387456, -1.71712542, -4.91501331, 0.1465137, 6.38665438},
                    }},
                },
            },
        };
        m[0] = m[1] = base;
        m[1].nonlinear_offset[0] = 0.0625f;
        m[1].nonlinear_offset[1] = 0.5f;
        m[1].nonlinear_offset[2] = 0.5f;
        m[1].nonlinear = (pl_matrix3x3) {{{1.0f, 0.0f, 1.5f}, {1.0f, -0.17f, -0.58f}, {1.0f, 1.88f, 0.0f}}};
        m[1].linear = (pl_matrix3x3) {{{0.95f, 0.04f, 0.01f}, {0.02f, 0.96f, 0.02f}, {0.0f, 0.03f, 0.97f}}};
        init = true;
    }
    return &m[variant ? 1 : 0];
}

API const void *shim_dovi_fixture_ptr(int variant) { return shim_dovi_fixture(variant); }

API int shim_run(struct shim_ctx *c, void *src_tex, void *dst_tex, int src_kind,
                 const float *sargs, const char *filter_name, const int *ops, const float *oargs,
                 float *out_peak /* avg_pq, max_pq or NULL */)
{
    pl_shader sh = pl_dispatch_begin(c->dp);
    pl_tex src = src_tex, dst = dst_tex;
    int ret = -1;

    /* sargs: [0..3] rect (all zero = full), [4] new_w, [5] new_h, [6] components, [7] scale,
     *        deband: [8] iterations [9] threshold [10] radius [11] grain [12..14] neutral
     *        scalers: [8] antiring (<0 = filter default); all: [15] address mode */
    struct pl_sample_src ssrc = {
        .tex = src,
        .rect = { sargs[0], sargs[1], sargs[2], sargs[3] },
        .new_w = (int) sargs[4],
        .new_h = (int) sargs[5],
        .components = (int) sargs[6],
        .scale = sargs[7],
        .address_mode = (int) sargs[15],    /* enum pl_tex_address_mode */
    };
    bool ok = true;
    switch (src_kind) {
    case 0: ok = pl_shader_sample_direct(sh, &ssrc); break;
    case 1: ok = pl_shader_sample_nearest(sh, &ssrc); break;
    case 2: ok = pl_shader_sample_bilinear(sh, &ssrc); break;
    case 3: {
        struct pl_deband_params dp = {
            .iterations = (int) sargs[8],
            .threshold = sargs[9],
            .radius = sargs[10],
            .grain = sargs[11],
            .grain_neutral = { sargs[12], sargs[13], sargs[14] },
        };
        pl_shader_deband(sh, &ssrc, &dp);
        break;
    }
    case 4:
    case 5: {
        const struct pl_filter_config *cfg = pl_find_filter_config(filter_name, 0);
        if (!cfg)
            goto done;
        struct pl_sample_filter_params fp = {
            .filter = *cfg,
            .lut = &c->objs.sampler,
        };
        if (sargs[8] >= 0)
            fp.antiring = sargs[8];
        ok = src_kind == 4 ? pl_shader_sample_polar(sh, &ssrc, &fp)
                           : pl_shader_sample_ortho2(sh, &ssrc, &fp);
        break;
    }
    case 6: ok = pl_shader_sample_bicubic(sh, &ssrc); break;
    case 7: ok = pl_shader_sample_hermite(sh, &ssrc); break;
    case 8: ok = pl_shader_sample_gaussian(sh, &ssrc); break;
    case 9: ok = pl_shader_sample_oversample(sh, &ssrc, sargs[8]); break; /* [8] threshold */
    default:
        goto done;
    }
    if (!ok || pl_shader_is_failed(sh))
        goto done;

    const float *a = oargs;
    for (; ops && *ops; ops++) {
        switch (*ops) {
        case 1:
        case 2: {
            struct pl_color_repr repr = {
                .sys = (int) a[0],
                .levels = (int) a[1],
                .bits = { (int) a[2], (int) a[3], (int) a[4] },
            };
            a += 5;
            if (*ops == 1)
                pl_shader_decode_color(sh, &repr, NULL);
            else
                pl_shader_encode_color(sh, &repr);
            break;
        }
        case 3:
        case 4: {
            struct pl_color_space csp = csp_from_flat(a);
            a += CSP_FLOATS;
            if (*ops == 3)
                pl_shader_linearize(sh, &csp);
            else
                pl_shader_delinearize(sh, &csp);
            break;
        }
        case 5: {
            struct pl_color_space s = csp_from_flat(a), d = csp_from_flat(a + CSP_FLOATS);
            a += 2 * CSP_FLOATS;
            struct pl_color_map_params par = a[0] ? pl_color_map_high_quality_params
                                                  : pl_color_map_default_params;
            par.contrast_recovery = 0;
            if (a[1] >= 0)
                par.tone_mapping_function = pl_tone_map_functions[(int) a[1]];
            if (a[2] >= 0)
                par.gamut_mapping = pl_gamut_map_functions[(int) a[2]];
            bool use_state = a[3];
            a += 4;
            pl_shader_color_map_ex(sh, &par, pl_color_map_args(
                .src = s,
                .dst = d,
                .state = use_state ? &c->objs.cmap : NULL,
            ));
            break;
        }
        case 6: {
            struct pl_dither_params dp = {
                .method = (int) a[1],
                .lut_size = (int) a[2],
                .temporal = a[3],
                .transfer = (int) a[4],
            };
            pl_shader_dither(sh, (int) a[0], &c->objs.dither, &dp);
            a += 5;
            break;
        }
        case 7:
        case 8: {
            struct pl_sigmoid_params sp = { a[0], a[1] };
            a += 2;
            if (*ops == 7)
                pl_shader_sigmoidize(sh, &sp);
            else
                pl_shader_unsigmoidize(sh, &sp);
            break;
        }
        case 9: {
            struct pl_color_space csp = csp_from_flat(a);
            bool hq = a[CSP_FLOATS];
            a += CSP_FLOATS + 1;
            if (!pl_shader_detect_peak(sh, csp, &c->objs.cmap,
                                       hq ? &pl_peak_detect_high_quality_params
                                          : &pl_peak_detect_default_params))
                goto done;
            break;
        }
        case 10: /* pl_shader_dovi_reshape with the reference's test fixture */
            pl_shader_dovi_reshape(sh, shim_dovi_fixture((int) a[0]));
            a += 1;
            break;
        case 11: { /* pl_shader_decode_color of a Dolby Vision repr (fixture, variant a[0]) */
            struct pl_color_repr repr = {
                .sys = PL_COLOR_SYSTEM_DOLBYVISION,
                .levels = PL_COLOR_LEVELS_FULL,
                .bits = { (int) a[1], (int) a[2], 0 },
                .dovi = shim_dovi_fixture((int) a[0]),
            };
            a += 3;
            pl_shader_decode_color(sh, &repr, NULL);
            break;
        }
        case 12: { /* pl_shader_extract_features */
            struct pl_color_space csp = csp_from_flat(a);
            a += CSP_FLOATS;
            pl_shader_extract_features(sh, csp);
            break;
        }
        default:
            goto done;
        }
    }
    if (pl_shader_is_failed(sh))
        goto done;

    ok = pl_dispatch_finish(c->dp, pl_dispatch_params( .shader = &sh, .target = dst ));
    if (ok && out_peak) {
        struct pl_hdr_metadata md = {0};
        if (pl_get_detected_hdr_metadata(c->objs.cmap, &md)) {
            out_peak[0] = md.avg_pq_y;
            out_peak[1] = md.max_pq_y;
        }
    }
    ret = ok ? 0 : -2;

done:
    pl_dispatch_abort(c->dp, &sh);
    return ret;
}

/* ---- custom LUTs ----------------------------------------------------------------------------
 * shim_custom_lut: sample_direct + pl_shader_custom_lut. data: RGB triples (red fastest);
 * size[1] == 0: 1-D; shaper_in / shaper_out: 9 floats (row major) or NULL. */
static struct pl_custom_lut make_lut(const float *data, const int size[3], const float *shaper_in,
                                     const float *shaper_out, uint64_t signature)
{
    struct pl_custom_lut lut = { .signature = signature, .size = { size[0], size[1], size[2] }, .data = data };
    for (int i = 0; i < 9; i++) {
        if (shaper_in)
            lut.shaper_in.m[i / 3][i % 3] = shaper_in[i];
        if (shaper_out)
            lut.shaper_out.m[i / 3][i % 3] = shaper_out[i];
    }
    return lut;
}

API int shim_custom_lut(struct shim_ctx *c, void *src_tex, void *dst_tex, const float *data,
                        const int *size, const float *shaper_in, const float *shaper_out,
                        unsigned long long signature)
{
    const struct pl_custom_lut lut = make_lut(data, size, shaper_in, shaper_out, signature);
    pl_shader sh = pl_dispatch_begin(c->dp);
    pl_shader_sample_direct(sh, pl_sample_src( .tex = src_tex ));
    pl_shader_custom_lut(sh, &lut, &c->lut_obj);
    return pl_dispatch_finish(c->dp, pl_dispatch_params( .shader = &sh, .target = dst_tex )) ? 0 : -1;
}

/* which: 0 pl_render_params.lut, 1 image.lut, 2 target.lut; type: enum pl_lut_type; the data
 * pointer must stay valid until the next shim_render returns. sys_in / sys_out: repr_in.sys /
 * repr_out.sys (frame LUT type guessing); trc_in/out, prim_in/out: color_in / color_out. */
API void shim_set_lut(struct shim_ctx *c, int which, int type, const float *data, const int *size,
                      const float *shaper_in, const float *shaper_out, unsigned long long signature,
                      const int *csp /* prim_in, trc_in, prim_out, trc_out, sys_in, sys_out */)
{
    c->lut[which] = make_lut(data, size, shaper_in, shaper_out, signature);
    c->lut_type[which] = type;
    if (csp) {
        c->lut[which].color_in.primaries = csp[0];
        c->lut[which].color_in.transfer = csp[1];
        c->lut[which].color_out.primaries = csp[2];
        c->lut[which].color_out.transfer = csp[3];
        c->lut[which].repr_in.sys = csp[4];
        c->lut[which].repr_out.sys = csp[5];
    }
}

/* pl_shader_deinterlace as a stand-alone dispatch: prev / next may be NULL */
API int shim_deinterlace(struct shim_ctx *c, void *cur, void *prev, void *next, void *dst, int field,
                         int first_field, int algo, int skip_spatial_check, int comp_mask)
{
    pl_shader sh = pl_dispatch_begin(c->dp);
    pl_shader_deinterlace(sh, pl_deinterlace_source(
        .cur.top = cur, .prev.top = prev, .next.top = next, .field = field,
        .first_field = first_field, .component_mask = comp_mask,
    ), pl_deinterlace_params( .algo = algo, .skip_spatial_check = skip_spatial_check ));
    return pl_dispatch_finish(c->dp, pl_dispatch_params( .shader = &sh, .target = dst )) ? 0 : -1;
}

/* The next shim_render call treats the image as an interlaced frame showing `field`, with the
 * given algorithm; prev / next: three plane textures each, or NULL. */
API void shim_set_deinterlace(struct shim_ctx *c, int field, int first_field, int algo, int skip,
                              void **prev, void **next)
{
    c->di_field = field;
    c->di_first = first_field;
    c->di_algo = algo;
    c->di_skip = skip;
    for (int k = 0; k < 3; k++) {
        c->di_prev[k] = prev ? prev[k] : NULL;
        c->di_next[k] = next ? next[k] : NULL;
    }
}

/* The next shim_render call becomes pl_render_image_mix: n frames of three planes each (planes =
 * n * 3 textures; the `planes` argument of shim_render is ignored), their timestamps relative to
 * the vsync, signatures, the vsync duration and the frame mixer ("" = none, i.e. nearest frame;
 * a filter preset name such as "oversample", "linear", "mitchell_clamp"). */
API void shim_set_mix(struct shim_ctx *c, int n, void **planes, const float *timestamps,
                      const unsigned long long *signatures, float vsync, const char *mixer)
{
    c->mix_n = n;
    for (int i = 0; i < n && i < 8; i++) {
        for (int k = 0; k < 3; k++)
            c->mix_planes[i][k] = planes[3 * i + k];
        c->mix_ts[i] = timestamps[i];
        c->mix_sig[i] = signatures[i];
    }
    c->mix_vsync = vsync;
    snprintf(c->mix_filter, sizeof(c->mix_filter), "%s", mixer ? mixer : "");
}

/* ---- pl_render_image driver ------------------------------------------------------------------
 * planes: up to 3 source textures (Y, U, V) or a single packed one (nplanes = 1, RGB)
 * img / tgt: repr (sys, levels, sample_depth, color_depth, bit_shift) followed by csp[20]
 * opts: [0] deband  [1] peak detect (0 off, 1 default, 2 hq)  [2] allow_delayed
 *       [3] dither (0 off, 1 blue noise default, 2 ordered LUT, 3 ordered fixed, 4 white)
 *       [4] colour map preset (0 default, 1 hq without contrast recovery)  [5] chroma location
 *       [6] sigmoid  [7] frame index (-1: renderer counts)  [8] error diffusion kernel idx+1
 *       [9] tone fn idx+1  [10] gamut fn idx+1 [11] temporal dither [12] film grain (1 = AV1 fixture, 2 = + overlap, 3 = H.274 fixture)
 *       [13] grain seed
 */
static struct pl_av1_grain_data shim_av1_fixture(void);

/* The next shim_render writes a planar target instead of `dst`: n = 3 one-component planes
 * (Y, Cb, Cr) or n = 2 (Y + interleaved CbCr, NV12 / P010 style); chroma planes may be
 * subsampled, sited at `chroma_loc`. */
API void shim_set_target_planes(struct shim_ctx *c, int n, void **texs, int chroma_loc)
{
    c->tgt_nplanes = n;
    c->tgt_chroma_loc = chroma_loc;
    for (int i = 0; i < n && i < 3; i++)
        c->tgt_planes[i] = texs[i];
}

API int shim_render(struct shim_ctx *c, int nplanes, void **planes, void *dst, const float *img,
                    const float *tgt, const int *opts, const char *upscaler,
                    const char *plane_upscaler, const char *downscaler, float *out_peak)
{
    if (!c->rr)
        c->rr = pl_renderer_create(shim_log, c->gpu);
    struct pl_frame image = {
        .num_planes = nplanes,
        .repr = { .sys = (int) img[0], .levels = (int) img[1],
                  .bits = { (int) img[2], (int) img[3], (int) img[4] } },
        .color = csp_from_flat(img + 5),
    };
    if (nplanes == 2) {
        /* semi-planar (NV12 / P010): Y + interleaved CbCr */
        image.planes[0] = (struct pl_plane) { .texture = planes[0], .components = 1,
                                              .component_mapping = { PL_CHANNEL_Y } };
        image.planes[1] = (struct pl_plane) { .texture = planes[1], .components = 2,
                                              .component_mapping = { PL_CHANNEL_CB, PL_CHANNEL_CR } };
        pl_frame_set_chroma_location(&image, opts[5] ? opts[5] : PL_CHROMA_LEFT);
    } else if (nplanes == 3) {
        for (int i = 0; i < 3; i++) {
            image.planes[i] = (struct pl_plane) {
                .texture = planes[i],
                .components = 1,
                .component_mapping = { i },
            };
        }
        pl_frame_set_chroma_location(&image, opts[5] ? opts[5] : PL_CHROMA_LEFT);
    } else {
        image.planes[0] = (struct pl_plane) {
            .texture = planes[0],
            .components = 3,
            .component_mapping = { 0, 1, 2 },
        };
    }
    if (opts[14]) { /* Dolby Vision: the reference's dovi_meta fixture, variant opts[14] - 1 */
        image.repr.sys = PL_COLOR_SYSTEM_DOLBYVISION;
        image.repr.dovi = shim_dovi_fixture(opts[14] - 1);
    }
    if (opts[12] == 3) {
        /* the reference's H.274 fixture (src/tests/utils.h:263-277): luma only */
        static const uint8_t lo = 10, hi = 250;
        static const int16_t vals[6] = { 16, 12, 14 };
        image.film_grain.type = PL_FILM_GRAIN_H274;
        image.film_grain.seed = (uint64_t) opts[13];
        image.film_grain.params.h274 = (struct pl_h274_grain_data) {
            .log2_scale_factor = 2,
            .component_model_present = { true },
            .num_intensity_intervals = { 1 },
            .num_model_values = { 3 },
            .intensity_interval_lower_bound = { &lo },
            .intensity_interval_upper_bound = { &hi },
            .comp_model_value = { &vals },
        };
    } else if (opts[12]) {
        image.film_grain.type = PL_FILM_GRAIN_AV1;
        image.film_grain.seed = (uint64_t) opts[13];
        image.film_grain.params.av1 = shim_av1_fixture();
        image.film_grain.params.av1.overlap = opts[12] > 1;
    }
    pl_tex dtex = dst;
    struct pl_frame target = {
        .num_planes = 1,
        .planes = {{
            .texture = dtex,
            .components = dtex->params.format->num_components,
            .component_mapping = { 0, 1, 2, 3 },
        }},
        .repr = { .sys = (int) tgt[0], .levels = (int) tgt[1],
                  .bits = { (int) tgt[2], (int) tgt[3], (int) tgt[4] } },
        .color = csp_from_flat(tgt + 5),
    };
    if (c->tgt_nplanes > 1) {
        target.num_planes = c->tgt_nplanes;
        target.planes[0] = (struct pl_plane) { .texture = c->tgt_planes[0], .components = 1,
                                               .component_mapping = { PL_CHANNEL_Y } };
        if (c->tgt_nplanes == 2) {
            target.planes[1] = (struct pl_plane) { .texture = c->tgt_planes[1], .components = 2,
                                                   .component_mapping = { PL_CHANNEL_CB, PL_CHANNEL_CR } };
        } else {
            for (int i = 1; i < 3; i++)
                target.planes[i] = (struct pl_plane) { .texture = c->tgt_planes[i], .components = 1,
                                                       .component_mapping = { i } };
        }
        pl_frame_set_chroma_location(&target, c->tgt_chroma_loc ? c->tgt_chroma_loc : PL_CHROMA_LEFT);
        c->tgt_nplanes = 0;
    }

    struct pl_deband_params deband = pl_deband_default_params;
    struct pl_peak_detect_params peak = opts[1] == 2 ? pl_peak_detect_high_quality_params
                                                     : pl_peak_detect_default_params;
    peak.allow_delayed = opts[2];
    struct pl_dither_params dither = pl_dither_default_params;
    if (opts[3] > 1)
        dither.method = opts[3] - 1;
    dither.temporal = opts[11];
    struct pl_color_map_params cmap = opts[4] ? pl_color_map_high_quality_params
                                              : pl_color_map_default_params;
    if (!opts[15]) /* opts[15]: keep the preset's contrast recovery (HQ: 0.3, smoothness 3.5) */
        cmap.contrast_recovery = 0;
    if (opts[9])
        cmap.tone_mapping_function = pl_tone_map_functions[opts[9] - 1];
    if (opts[10])
        cmap.gamut_mapping = pl_gamut_map_functions[opts[10] - 1];
    struct pl_render_params par = {
        PL_RENDER_DEFAULTS
        .upscaler = upscaler && *upscaler ? pl_find_filter_config(upscaler, 0) : NULL,
        .downscaler = downscaler && *downscaler ? pl_find_filter_config(downscaler, 0) : NULL,
        .plane_upscaler = plane_upscaler && *plane_upscaler ? pl_find_filter_config(plane_upscaler, 0) : NULL,
        .deband_params = opts[0] ? &deband : NULL,
        .peak_detect_params = opts[1] ? &peak : NULL,
        .dither_params = opts[3] ? &dither : NULL,
        .sigmoid_params = opts[6] ? &pl_sigmoid_default_params : NULL,
        .error_diffusion = opts[8] ? pl_error_diffusion_kernels[opts[8] - 1] : NULL,
    };
    par.color_map_params = &cmap;
    if (c->lut[0].data) {
        par.lut = &c->lut[0];
        par.lut_type = c->lut_type[0];
    }
    if (c->lut[1].data) {
        image.lut = &c->lut[1];
        image.lut_type = c->lut_type[1];
    }
    if (c->lut[2].data) {
        target.lut = &c->lut[2];
        target.lut_type = c->lut_type[2];
    }
    if (opts[7] >= 0)
        pl_renderer_set_frame_index(c->rr, (uint8_t) opts[7]);
    struct pl_deinterlace_params dipar = { .algo = c->di_algo, .skip_spatial_check = c->di_skip };
    struct pl_frame di_prev = image, di_next = image;
    if (c->di_field) {
        image.field = c->di_field;
        image.first_field = c->di_first;
        par.deinterlace_params = &dipar;
        if (c->di_prev[0]) {
            for (int k = 0; k < 3; k++)
                di_prev.planes[k].texture = c->di_prev[k];
            image.prev = &di_prev;
        }
        if (c->di_next[0]) {
            for (int k = 0; k < 3; k++)
                di_next.planes[k].texture = c->di_next[k];
            image.next = &di_next;
        }
        c->di_field = 0;
    }
    bool rendered;
    if (c->mix_n) {
        struct pl_frame frames[8];
        const struct pl_frame *fptr[8];
        for (int i = 0; i < c->mix_n; i++) {
            frames[i] = image;
            for (int k = 0; k < 3; k++)
                frames[i].planes[k].texture = c->mix_planes[i][k];
            fptr[i] = &frames[i];
        }
        const struct pl_frame_mix mix = {
            .num_frames = c->mix_n,
            .frames = fptr,
            .signatures = c->mix_sig,
            .timestamps = c->mix_ts,
            .vsync_duration = c->mix_vsync,
        };
        par.frame_mixer = c->mix_filter[0] ? pl_find_filter_config(c->mix_filter, 0) : NULL;
        rendered = pl_render_image_mix(c->rr, &mix, &target, &par);
        c->mix_n = 0;
    } else {
        rendered = pl_render_image(c->rr, &image, &target, &par);
    }
    memset(c->lut, 0, sizeof(c->lut));
    if (!rendered)
        return -1;
    if (out_peak) {
        struct pl_hdr_metadata md = {0};
        if (pl_renderer_get_hdr_metadata(c->rr, &md)) {
            out_peak[0] = md.avg_pq_y;
            out_peak[1] = md.max_pq_y;
        }
    }
    return pl_renderer_get_errors(c->rr).errors;
}

API void shim_renderer_reset(struct shim_ctx *c)
{
    pl_renderer_destroy(&c->rr);
    pl_shader_obj_destroy(&c->objs.cmap); /* forget the detected peak / cached LUTs */
}

/* pl_shader_error_diffusion as a stand-alone compute dispatch (src/tests/gpu_tests.c:1003-1040) */
API int shim_error_diffusion(struct shim_ctx *c, void *src_tex, void *dst_tex, int new_depth,
                             const char *kernel)
{
    pl_shader sh = pl_dispatch_begin(c->dp);
    bool ok = pl_shader_error_diffusion(sh, pl_error_diffusion_params(
        .input_tex = src_tex,
        .output_tex = dst_tex,
        .new_depth = new_depth,
        .kernel = pl_find_error_diffusion_kernel(kernel),
    ));
    if (!ok) {
        pl_dispatch_abort(c->dp, &sh);
        return -1;
    }
    return pl_dispatch_compute(c->dp, pl_dispatch_compute_params(
        .shader = &sh,
        .dispatch_size = {1, 1, 1},
    )) ? 0 : -2;
}

/* host tables of the AV1 grain for the reference's fixture (no GPU) */
static struct pl_av1_grain_data shim_av1_fixture(void);
API int shim_av1_tables(unsigned seed, int color_depth, int levels, int sys, int sub_x, int sub_y,
                        int tex_w, int tex_h, float *gy, float *gcb, float *gcr, uint32_t *offsets,
                        float *scaling)
{
    struct pl_film_grain_data data = {
        .type = PL_FILM_GRAIN_AV1,
        .seed = seed,
        .params.av1 = shim_av1_fixture(),
    };
    struct pl_color_repr repr = { .sys = sys, .levels = levels, .bits = { 16, color_depth, 0 } };
    return pl_cuda_av1_tables_host(&data, &repr, sub_x, sub_y, tex_w, tex_h, gy, gcb, gcr, offsets,
                                   scaling) ? 0 : -1;
}

API const void *shim_av1_fixture_ptr(void)
{
    static struct pl_av1_grain_data d;
    d = shim_av1_fixture();
    return &d;
}

/* pl_shader_film_grain (AV1, reference fixture) on one plane -> dst (src/tests/gpu_tests.c:787-835) */
API int shim_film_grain(struct shim_ctx *c, void *tex, void *luma_tex, void *dst, int chroma,
                        unsigned seed, int overlap, int color_depth, int levels, int sys)
{
    struct pl_color_repr repr = { .sys = sys, .levels = levels, .bits = { 16, color_depth, 0 } };
    struct pl_film_grain_params par = {
        .data = { .type = PL_FILM_GRAIN_AV1, .seed = seed, .params.av1 = shim_av1_fixture() },
        .tex = tex,
        .luma_tex = luma_tex,
        .repr = &repr,
        .components = chroma ? 2 : 1,
        .component_mapping = { chroma ? PL_CHANNEL_CB : PL_CHANNEL_Y, PL_CHANNEL_CR },
        .luma_comp = 0,
    };
    par.data.params.av1.overlap = overlap;
    pl_shader sh = pl_dispatch_begin(c->dp);
    if (!pl_shader_film_grain(sh, &c->objs.grain[chroma ? 1 : 0], &par)) {
        pl_dispatch_abort(c->dp, &sh);
        return -1;
    }
    return pl_dispatch_finish(c->dp, pl_dispatch_params( .shader = &sh, .target = dst )) ? 0 : -2;
}

/* pl_shader_film_grain (H.274) on one plane -> dst. The model arrives flat: per channel up to 8
 * intensity intervals, `lower`/`upper` [3][8], `values` [3][8][6]. */
API int shim_film_grain_h274(struct shim_ctx *c, void *tex, void *dst, int comps, const int *map,
                             unsigned seed, int color_depth, int levels, int sys, int log2_scale,
                             const int *present, const int *nint, const int *nvals,
                             const uint8_t *lower, const uint8_t *upper, const int16_t *values,
                             int state_slot)
{
    struct pl_color_repr repr = { .sys = sys, .levels = levels, .bits = { 16, color_depth, 0 } };
    struct pl_film_grain_params par = {
        .data = { .type = PL_FILM_GRAIN_H274, .seed = seed },
        .tex = tex,
        .repr = &repr,
        .components = comps,
    };
    struct pl_h274_grain_data *d = &par.data.params.h274;
    d->log2_scale_factor = log2_scale;
    for (int ch = 0; ch < 3; ch++) {
        d->component_model_present[ch] = present[ch];
        d->num_intensity_intervals[ch] = nint[ch];
        d->num_model_values[ch] = nvals[ch];
        d->intensity_interval_lower_bound[ch] = lower + 8 * ch;
        d->intensity_interval_upper_bound[ch] = upper + 8 * ch;
        d->comp_model_value[ch] = (const int16_t (*)[6]) (values + 48 * ch);
    }
    for (int i = 0; i < comps; i++)
        par.component_mapping[i] = map[i];
    if (!pl_needs_film_grain(&par))
        return 1;
    pl_shader sh = pl_dispatch_begin(c->dp);
    if (!pl_shader_film_grain(sh, &c->objs.grain[state_slot & 1], &par)) {
        pl_dispatch_abort(c->dp, &sh);
        return -1;
    }
    return pl_dispatch_finish(c->dp, pl_dispatch_params( .shader = &sh, .target = dst )) ? 0 : -2;
}

/* utils/upload.h: pl_upload_plane (via_buf: through a pl_buf instead of host memory) or
 * pl_recreate_plane (pixels == NULL). layout = size[4], pad[4], map[4]; info out = components,
 * component_mapping[4]. Returns the texture (destroy with shim_tex_destroy) or NULL. */
API void *shim_upload_plane(struct shim_ctx *c, int type, const int *layout, int pixel_stride,
                            int row_stride, int swapped, int w, int h, const void *pixels,
                            size_t nbytes, int via_buf, int *info)
{
    struct pl_plane_data d = {
        .type = type, .width = w, .height = h, .pixel_stride = pixel_stride,
        .row_stride = row_stride, .swapped = swapped,
    };
    for (int i = 0; i < 4; i++) {
        d.component_size[i] = layout[i];
        d.component_pad[i] = layout[4 + i];
        d.component_map[i] = layout[8 + i];
    }
    struct pl_plane plane = {0};
    pl_tex tex = NULL;
    pl_buf buf = NULL;
    bool ok;
    if (!pixels) {
        ok = pl_recreate_plane(c->gpu, &plane, &tex, &d);
    } else if (via_buf) {
        /* offset the data inside the buffer to exercise buf_offset */
        buf = pl_buf_create(c->gpu, pl_buf_params( .size = nbytes + 64, .host_writable = true,
                                                   .storable = via_buf > 1 ));
        if (!buf)
            return NULL;
        pl_buf_write(c->gpu, buf, 64, pixels, nbytes);
        d.buf = buf;
        d.buf_offset = 64;
        ok = pl_upload_plane(c->gpu, &plane, &tex, &d);
    } else {
        d.pixels = pixels;
        ok = pl_upload_plane(c->gpu, &plane, &tex, &d);
    }
    pl_gpu_finish(c->gpu);
    pl_buf_destroy(c->gpu, &buf);
    if (!ok) {
        pl_tex_destroy(c->gpu, &tex);
        return NULL;
    }
    info[0] = plane.components;
    for (int i = 0; i < 4; i++)
        info[1 + i] = plane.component_mapping[i];
    info[5] = tex->params.format->num_components;
    info[6] = (int) tex->params.format->texel_size;
    info[7] = tex->params.renderable;
    return (void *) tex;
}

/* host-only: the H.274 grain data base (no GPU needed) */
API int shim_h274_database(float *out)
{
    return pl_cuda_h274_database_host(out) ? 0 : -1;
}

API int shim_math_eval(struct shim_ctx *c, int fn, const float *a, const float *b, float *out,
                       size_t n)
{
    return pl_cuda_math_eval(c->gpu, fn, a, b, out, n) ? 0 : -1;
}

/* the AV1 grain parameters of the reference's own test fixture (src/tests/utils.h:240-262) */
static struct pl_av1_grain_data shim_av1_fixture(void)
{
    return (struct pl_av1_grain_data) {
        .num_points_y = 6,
        .points_y = {{0, 4}, {27, 33}, {54, 55}, {67, 61}, {108, 71}, {255, 72}},
        .chroma_scaling_from_luma = false,
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
        .grain_scale_shift = 0,
        .uv_mult = {0, 0},
        .uv_mult_luma = {64, 64},
        .uv_offset = {0, 0},
    };
}

/* host-only: the peak recurrence of the frame-parallel scheme (no GPU needed) */
API int shim_peak_update(float state[2], const void *buf, int hq)
{
    struct pl_peak_state st = { .avg_pq = state[0], .max_pq = state[1] };
    const struct pl_peak_detect_params *par = hq ? &pl_peak_detect_high_quality_params
                                                 : &pl_peak_detect_default_params;
    bool used = pl_peak_state_update(&st, buf, par);
    state[0] = st.avg_pq;
    state[1] = st.max_pq;
    return used;
}
