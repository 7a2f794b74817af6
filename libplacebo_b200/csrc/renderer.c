// pl_renderer on the CUDA backend: the per-frame planner.
//
// Restates the planning logic of the reference's src/renderer.c for the hot path
// (pl_render_image :3433 -> pass_read_image :1553, pass_scale_main :1964, hdr_update_peak
// :1183, pass_convert_colors :2157, pass_output_target :2586), building pl_shader op lists
// instead of GLSL and materialising intermediates in the SAME places and formats (FBO format
// selection find_fbo_format :383-434, img_tex :505-547), because every intermediate store is
// a rounding step that is part of the reference's result:
//     A. same-sized chroma planes are merged into one 2-component texture   (:1617-1699)
//     B. deband -> [film grain] per plane; planes that need resampling are stored to an
//        fp16 FBO first, then scaled by the plane scaler                    (:1763-1871)
//     C. all planes are merged 1:1 into the main pass, decoded to RGB       (:1874-1957)
//     D. main scaler (with optional linearisation / sigmoidisation), HDR peak detection,
//        colour mapping, encoding, dithering, store                         (:1964-2964)
// Features of pl_render_params outside the hot path (hooks, custom LUTs, ICC, overlays,
// distortion, blending, deinterlacing, rotation, frame mixing) are rejected loudly instead of
// being silently ignored.
#define _GNU_SOURCE
#include "shader_priv.h"

#include <math.h>

bool plb_dispatch_run(pl_dispatch dp, pl_shader sh, pl_tex target, pl_rect2d rc, int w, int h,
                      pl_timer timer);

#define MAX_FBOS 16

struct sampler_state {
    pl_shader_obj up, down;
};

struct pl_renderer_t {
    pl_log log;
    pl_gpu gpu;
    pl_dispatch dp;
    enum pl_render_error errors;
    pl_fmt fbofmt[5];
    pl_tex fbos[MAX_FBOS];
    bool fbo_used[MAX_FBOS];
    struct sampler_state sampler_main, sampler_contrast, samplers_src[4], samplers_dst[4];
    pl_shader_obj tone_map_state, dither_state, grain_state[4];
    pl_shader_obj lut_state[3];     // LUT_IMAGE, LUT_TARGET, LUT_PARAMS (renderer.c:96-100)
    // frame mixing cache (renderer.c:73-95 `struct cached_frame`): frames rendered up to the
    // target colour space, keyed by the caller's signature
    struct cached_frame {
        uint64_t signature;
        uint64_t params_hash;
        struct pl_color_space color;
        pl_rect2df crop;
        int comps;
        pl_tex tex;
        bool evict;
    } frames[16];
    int num_frames;
    int prev_dither;
    bool have_frame_index;
    uint8_t frame_index;
    bool have_peak_override;
    struct pl_peak_state peak_override;
    // planar pair views created for the current frame (freed when the frame has been recorded)
    pl_tex views[4];
    int num_views;
    // info_callback plumbing (renderer.c:3320-3345 `info_callback`)
    const struct pl_render_params *info_params;
    int info_index;
    // Frame-parallel mode: pl_renderer_measure_peak materialises the merged source pass (the
    // rgba16hf image the sequential same-frame flow would also produce) and the very next
    // pl_render_image of the same planes continues from it instead of redoing the pass.
    // Several frames may be measured ahead of their render (the multi-GPU pipeline keeps the next
    // frame's measurement in flight while the statistics of the current one are exchanged).
    struct measured_image {
        bool valid;
        uint64_t seq;               // measurement counter when stored (stale entries expire)
        pl_tex tex;                 // one of rr->fbos, kept reserved while valid
        pl_tex planes[4];
        // everything that determines the stored image: a later pl_render_image only continues
        // from it when all of this matches
        struct pl_render_params params;
        const struct pl_render_params *params_ptr;
        pl_rect2df crop;
        struct pl_color_repr in_repr;
        struct pl_color_space in_color;
        struct pl_film_grain_data grain;
        uint8_t frame_index;
        int w, h, comps;
        pl_rect2df rect, ref_rect;
        struct pl_color_repr repr;
        struct pl_color_space color;
    } measured[4];
    uint64_t measure_seq;
};

static void free_views(pl_renderer rr)
{
    for (int i = 0; i < rr->num_views; i++)
        pl_tex_destroy(rr->gpu, &rr->views[i]);
    rr->num_views = 0;
}

static void renderer_info_cb(void *priv, const struct pl_dispatch_info *dinfo)
{
    pl_renderer rr = priv;
    if (!rr->info_params || !rr->info_params->info_callback)
        return;
    rr->info_params->info_callback(rr->info_params->info_priv, &(struct pl_render_info) {
        .pass = dinfo,
        .stage = PL_RENDER_STAGE_FRAME,
        .index = rr->info_index++,
    });
}

const struct pl_render_params pl_render_fast_params = { PL_RENDER_DEFAULTS };
const struct pl_render_params pl_render_default_params = {
    PL_RENDER_DEFAULTS
    .upscaler           = &pl_filter_lanczos,
    .downscaler         = &pl_filter_hermite,
    .sigmoid_params     = &pl_sigmoid_default_params,
    .dither_params      = &pl_dither_default_params,
    .peak_detect_params = &pl_peak_detect_default_params,
};
const struct pl_render_params pl_render_high_quality_params = {
    PL_RENDER_DEFAULTS
    .upscaler           = &pl_filter_ewa_lanczossharp,
    .downscaler         = &pl_filter_hermite,
    .sigmoid_params     = &pl_sigmoid_default_params,
    .peak_detect_params = &pl_peak_detect_high_quality_params,
    .color_map_params   = &pl_color_map_high_quality_params,
    .dither_params      = &pl_dither_default_params,
    .deband_params      = &pl_deband_default_params,
};

pl_renderer pl_renderer_create(pl_log log, pl_gpu gpu)
{
    if (!gpu) {
        plb_msg(log, PL_LOG_ERR, "pl_renderer_create: no gpu");
        return NULL;
    }
    struct pl_renderer_t *rr = calloc(1, sizeof(*rr));
    rr->log = log;
    rr->gpu = gpu;
    rr->dp = pl_dispatch_create(log, gpu);
    return rr;
}

void pl_renderer_destroy(pl_renderer *prr)
{
    pl_renderer rr = prr ? *prr : NULL;
    if (!rr)
        return;
    for (int i = 0; i < MAX_FBOS; i++)
        pl_tex_destroy(rr->gpu, &rr->fbos[i]);
    pl_shader_obj_destroy(&rr->sampler_main.up);
    pl_shader_obj_destroy(&rr->sampler_main.down);
    pl_shader_obj_destroy(&rr->sampler_contrast.up);
    pl_shader_obj_destroy(&rr->sampler_contrast.down);
    for (int i = 0; i < 3; i++)
        pl_shader_obj_destroy(&rr->lut_state[i]);
    for (int i = 0; i < rr->num_frames; i++)
        pl_tex_destroy(rr->gpu, &rr->frames[i].tex);
    for (int i = 0; i < 4; i++) {
        pl_shader_obj_destroy(&rr->samplers_dst[i].up);
        pl_shader_obj_destroy(&rr->samplers_dst[i].down);
    }
    for (int i = 0; i < 4; i++) {
        pl_shader_obj_destroy(&rr->samplers_src[i].up);
        pl_shader_obj_destroy(&rr->samplers_src[i].down);
        pl_shader_obj_destroy(&rr->grain_state[i]);
    }
    pl_shader_obj_destroy(&rr->tone_map_state);
    pl_shader_obj_destroy(&rr->dither_state);
    pl_dispatch_destroy(&rr->dp);
    free(rr);
    *prr = NULL;
}

struct pl_render_errors pl_renderer_get_errors(pl_renderer rr)
{
    return (struct pl_render_errors) { .errors = rr->errors };
}

void pl_renderer_reset_errors(pl_renderer rr, const struct pl_render_errors *errors)
{
    if (!errors) {
        rr->errors = PL_RENDER_ERR_NONE;
        return;
    }
    rr->errors &= ~errors->errors;
}

void pl_renderer_flush_cache(pl_renderer rr)
{
    pl_reset_detected_peak(rr->tone_map_state);
}

bool pl_renderer_get_hdr_metadata(pl_renderer rr, struct pl_hdr_metadata *metadata)
{
    return pl_get_detected_hdr_metadata(rr->tone_map_state, metadata);
}

void pl_renderer_set_frame_index(pl_renderer rr, uint8_t index)
{
    rr->have_frame_index = true;
    rr->frame_index = index;
}

void pl_renderer_set_peak_state(pl_renderer rr, const struct pl_peak_state *state)
{
    rr->have_peak_override = state != NULL;
    if (state)
        rr->peak_override = *state;
}

// ---------------------------------------------------------------------------------------
// frame fix-ups (renderer.c:3043-3233)
// ---------------------------------------------------------------------------------------

enum plane_type { PLANE_INVALID = 0, PLANE_ALPHA, PLANE_CHROMA, PLANE_LUMA, PLANE_RGB, PLANE_XYZ };

static enum plane_type detect_plane_type(const struct pl_plane *plane,
                                         const struct pl_color_repr *repr)
{
    if (pl_color_system_is_ycbcr_like(repr->sys)) {
        int t = PLANE_INVALID;
        for (int c = 0; c < plane->components; c++) {
            switch (plane->component_mapping[c]) {
            case PL_CHANNEL_Y: t = PLB_MAX(t, PLANE_LUMA); break;
            case PL_CHANNEL_A: t = PLB_MAX(t, PLANE_ALPHA); break;
            case PL_CHANNEL_CB:
            case PL_CHANNEL_CR: t = PLB_MAX(t, PLANE_CHROMA); break;
            default: break;
            }
        }
        return t;
    }
    if (plane->components == 1 && plane->component_mapping[0] == PL_CHANNEL_A)
        return PLANE_ALPHA;
    return repr->sys == PL_COLOR_SYSTEM_XYZ ? PLANE_XYZ : PLANE_RGB;
}

// index of the plane the frame's geometry refers to: the first one carrying luma / RGB / XYZ
static int frame_ref(const struct pl_frame *frame)
{
    int ref = 0;
    for (int i = frame->num_planes - 1; i >= 0; i--) {
        const enum plane_type t = detect_plane_type(&frame->planes[i], &frame->repr);
        if (t == PLANE_RGB || t == PLANE_LUMA || t == PLANE_XYZ)
            ref = i;
    }
    return ref;
}

static enum pl_color_primaries primaries_guess(int width, int height)
{
    if (width >= 1280 || height > 576)
        return PL_COLOR_PRIM_BT_709;
    switch (height) {
    case 576: return PL_COLOR_PRIM_BT_601_625;
    case 480:
    case 486: return PL_COLOR_PRIM_BT_601_525;
    default:  return PL_COLOR_PRIM_BT_709;
    }
}

// Fill in what a caller may leave open (fix_frame, renderer.c:3122-3160): XYZ implies DCI-P3 /
// ST 428, unknown primaries are guessed from the resolution, alpha only exists if a plane maps
// it, and UNORM textures define the sample depth (colour depth at most that, rest = shift).
static void fix_frame(struct pl_frame *frame)
{
    const pl_tex ref = frame->planes[frame_ref(frame)].texture;
    struct pl_color_space *csp = &frame->color;
    if (frame->repr.sys == PL_COLOR_SYSTEM_XYZ) {
        csp->primaries = PL_COLOR_PRIM_DCI_P3;
        csp->transfer = PL_COLOR_TRC_ST428;
    } else if (ref && !csp->primaries) {
        csp->primaries = primaries_guess(ref->params.w, ref->params.h);
    }

    int alpha_planes = 0;
    for (int p = 0; p < frame->num_planes; p++) {
        const struct pl_plane *pl = &frame->planes[p];
        for (int c = 0; c < pl->components; c++)
            alpha_planes += pl->component_mapping[c] == PL_CHANNEL_A;
    }
    if (!alpha_planes)
        frame->repr.alpha = PL_ALPHA_NONE;

    struct pl_bit_encoding *be = &frame->repr.bits;
    if (be->sample_depth || !ref || ref->params.format->type != PL_FMT_UNORM)
        return;
    const int depth = ref->params.format->component_depth[0];
    const int color = be->color_depth ? PLB_MIN(be->color_depth, depth) : depth;
    be->sample_depth = depth;
    be->color_depth = color;
    be->bit_shift += depth - color;
}

struct pass_state {
    pl_renderer rr;
    const struct pl_render_params *params;
    struct pl_frame image, target;
    int src_ref, dst_ref;
    pl_rect2df ref_rect;
    pl_rect2d dst_rect;
    bool need_peak_fbo;
    bool from_measured;    // pass->tex is the image pl_renderer_measure_peak stored (already an FBO)
    // current image
    pl_shader sh;       // pending shader (NULL if materialised in `tex`)
    pl_tex tex;
    int w, h, comps;
    pl_rect2df rect;
    struct pl_color_repr repr;
    struct pl_color_space color;
};

enum { LUT_IMAGE, LUT_TARGET, LUT_PARAMS };

// renderer.c:1447-1469
static enum pl_lut_type guess_frame_lut_type(const struct pl_frame *frame, bool reversed)
{
    if (!frame->lut)
        return PL_LUT_UNKNOWN;
    if (frame->lut_type)
        return frame->lut_type;
    enum pl_color_system sys_in = frame->lut->repr_in.sys, sys_out = frame->lut->repr_out.sys;
    if (reversed) {
        const enum pl_color_system t = sys_in;
        sys_in = sys_out;
        sys_out = t;
    }
    if (sys_in == PL_COLOR_SYSTEM_RGB && sys_out == sys_in)
        return PL_LUT_NORMALIZED;
    if (sys_in == frame->repr.sys && sys_out == PL_COLOR_SYSTEM_RGB)
        return PL_LUT_CONVERSION;
    return PL_LUT_NATIVE; // unknown: the default
}

static bool fix_refs_and_rects(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    struct pl_frame *image = &pass->image, *target = &pass->target;
    pl_rect2df *dst = &target->crop, *src = &image->crop;
    pass->dst_ref = frame_ref(target);
    pl_tex dst_ref = target->planes[pass->dst_ref].texture;
    const int dst_w = dst_ref->params.w, dst_h = dst_ref->params.h;
    if ((!dst->x0 && !dst->x1) || (!dst->y0 && !dst->y1)) {
        dst->x0 = dst->y0 = 0;
        dst->x1 = dst_w;
        dst->y1 = dst_h;
    }
    pass->src_ref = frame_ref(image);
    pl_tex src_ref = image->planes[pass->src_ref].texture;
    if ((!src->x0 && !src->x1) || (!src->y0 && !src->y1)) {
        src->x0 = src->y0 = 0;
        src->x1 = src_ref->params.w;
        src->y1 = src_ref->params.h;
    }
    if (image->rotation != target->rotation) {
        PLB_ERR(rr->log, "pl_render_image: rotation is outside the CUDA hot path");
        return false;
    }
    if (src->x0 > src->x1 || src->y0 > src->y1 || dst->x0 > dst->x1 || dst->y0 > dst->y1) {
        PLB_ERR(rr->log, "pl_render_image: flipped crops are outside the CUDA hot path");
        return false;
    }
    // snap the destination to whole pixels and adjust the source accordingly (:3089-3120)
    const float rx0 = roundf(fminf(fmaxf(dst->x0, 0.0), dst_w)),
                ry0 = roundf(fminf(fmaxf(dst->y0, 0.0), dst_h)),
                rx1 = roundf(fminf(fmaxf(dst->x1, 0.0), dst_w)),
                ry1 = roundf(fminf(fmaxf(dst->y1, 0.0), dst_h));
    const float scale_x = pl_rect_w(*src) / pl_rect_w(*dst),
                scale_y = pl_rect_h(*src) / pl_rect_h(*dst),
                base_x = src->x0, base_y = src->y0;
    src->x0 = base_x + (rx0 - dst->x0) * scale_x;
    src->x1 = base_x + (rx1 - dst->x0) * scale_x;
    src->y0 = base_y + (ry0 - dst->y0) * scale_y;
    src->y1 = base_y + (ry1 - dst->y0) * scale_y;
    *dst = (pl_rect2df) { rx0, ry0, rx1, ry1 };
    pass->ref_rect = *src;
    pass->dst_rect = (pl_rect2d) { dst->x0, dst->y0, dst->x1, dst->y1 };
    return true;
}

static bool pass_fix_frames(struct pass_state *pass)
{
    struct pl_frame *image = &pass->image, *target = &pass->target;
    if (!fix_refs_and_rects(pass))
        return false;
    fix_frame(image);
    pl_color_space_infer_map(&image->color, &target->color);
    fix_frame(target);
    if (image->repr.alpha == PL_ALPHA_UNKNOWN)
        image->repr.alpha = PL_ALPHA_INDEPENDENT;
    if (target->repr.alpha == PL_ALPHA_UNKNOWN)
        target->repr.alpha = PL_ALPHA_PREMULTIPLIED;
    return true;
}

void pl_frames_infer(pl_renderer rr, struct pl_frame *image, struct pl_frame *target)
{
    struct pass_state pass = { .rr = rr, .image = *image, .target = *target };
    if (pass_fix_frames(&pass)) {
        *image = pass.image;
        *target = pass.target;
    }
}

// renderer.h:849-854: the same inference for a frame mix, based on the frame nearest to the vsync
void pl_frames_infer_mix(pl_renderer rr, const struct pl_frame_mix *mix, struct pl_frame *target,
                         struct pl_frame *out_ref)
{
    const struct pl_frame *nearest = mix ? pl_frame_mix_nearest(mix) : NULL;
    if (!nearest || !target)
        return; // nothing to infer from (a mix without frames is outside this backend)
    struct pl_frame image = *nearest;
    pl_frames_infer(rr, &image, target);
    if (out_ref)
        *out_ref = image;
}

void pl_frame_set_chroma_location(struct pl_frame *frame, enum pl_chroma_location loc)
{
    pl_tex ref = frame->planes[frame_ref(frame)].texture;
    for (int i = 0; i < frame->num_planes; i++) {
        struct pl_plane *plane = &frame->planes[i];
        bool apply;
        if (ref) {
            apply = plane->texture->params.w < ref->params.w ||
                    plane->texture->params.h < ref->params.h;
        } else {
            apply = detect_plane_type(plane, &frame->repr) == PLANE_CHROMA;
        }
        if (apply)
            pl_chroma_location_offset(loc, &plane->shift_x, &plane->shift_y);
    }
}

bool pl_frame_is_cropped(const struct pl_frame *frame)
{
    int x0 = roundf(PLB_MIN(frame->crop.x0, frame->crop.x1)),
        y0 = roundf(PLB_MIN(frame->crop.y0, frame->crop.y1)),
        x1 = roundf(PLB_MAX(frame->crop.x0, frame->crop.x1)),
        y1 = roundf(PLB_MAX(frame->crop.y0, frame->crop.y1));
    pl_tex ref = frame->planes[frame_ref(frame)].texture;
    if (!x0 && !x1) x1 = ref->params.w;
    if (!y0 && !y1) y1 = ref->params.h;
    return x0 > 0 || y0 > 0 || x1 < ref->params.w || y1 < ref->params.h;
}

// ---------------------------------------------------------------------------------------
// FBOs (renderer.c:383-547)
// ---------------------------------------------------------------------------------------

static bool find_fbo_format(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    if (params->disable_fbos || (rr->errors & PL_RENDER_ERR_FBO)) {
        PLB_ERR(rr->log, "pl_render_image: rendering without FBOs is not supported on the "
                "CUDA backend");
        return false;
    }
    static const struct { enum pl_fmt_type type; int depth; } prefs[] = {
        { PL_FMT_FLOAT, 16 }, { PL_FMT_UNORM, 16 }, { PL_FMT_UNORM, 8 },
    };
    for (size_t i = 0; i < PLB_ARRAY_SIZE(prefs); i++) {
        if (params->force_low_bit_depth_fbos && prefs[i].depth > 8)
            continue;
        pl_fmt fmt = pl_find_fmt(rr->gpu, prefs[i].type, 4, prefs[i].depth, 0,
                                 PL_FMT_CAP_RENDERABLE | PL_FMT_CAP_LINEAR);
        if (!fmt)
            continue;
        rr->fbofmt[4] = fmt;
        for (int c = 3; c >= 1; c--) {
            rr->fbofmt[c] = pl_find_fmt(rr->gpu, prefs[i].type, c, prefs[i].depth, 0, fmt->caps);
            rr->fbofmt[c] = PLB_DEF(rr->fbofmt[c], rr->fbofmt[c + 1]);
        }
        return true;
    }
    return false;
}

static pl_tex get_fbo(struct pass_state *pass, int w, int h, pl_fmt fmt, int comps)
{
    pl_renderer rr = pass->rr;
    fmt = PLB_DEF(fmt, rr->fbofmt[PLB_DEF(comps, 4)]);
    int free_slot = -1;
    for (int i = 0; i < MAX_FBOS; i++) {
        if (rr->fbo_used[i])
            continue;
        pl_tex t = rr->fbos[i];
        if (t && t->params.w == w && t->params.h == h && t->params.format == fmt) {
            rr->fbo_used[i] = true;
            return t;
        }
        if (free_slot < 0 || (!t && rr->fbos[free_slot]))
            free_slot = i;
    }
    if (free_slot < 0) {
        PLB_ERR(rr->log, "out of FBO slots");
        return NULL;
    }
    if (!pl_tex_recreate(rr->gpu, &rr->fbos[free_slot], pl_tex_params(
            .w = w, .h = h, .format = fmt,
            .sampleable = true, .renderable = true, .storable = true)))
        return NULL;
    rr->fbo_used[free_slot] = true;
    return rr->fbos[free_slot];
}

// img_tex: materialise a pending shader
static pl_tex shader_to_tex(struct pass_state *pass, pl_shader *sh, int w, int h, pl_fmt fmt,
                            int comps)
{
    pl_renderer rr = pass->rr;
    pl_tex tex = get_fbo(pass, w, h, fmt, comps);
    if (!tex) {
        pl_dispatch_abort(rr->dp, sh);
        rr->errors |= PL_RENDER_ERR_FBO;
        return NULL;
    }
    if (!pl_dispatch_finish(rr->dp, pl_dispatch_params( .shader = sh, .target = tex ))) {
        PLB_ERR(rr->log, "Failed dispatching intermediate pass!");
        return NULL;
    }
    return tex;
}

static pl_shader begin(struct pass_state *pass)
{
    return pl_dispatch_begin(pass->rr->dp);
}

static pl_tex img_tex(struct pass_state *pass)
{
    if (pass->tex)
        return pass->tex;
    pass->tex = shader_to_tex(pass, &pass->sh, pass->w, pass->h, NULL, pass->comps);
    pass->sh = NULL;
    return pass->tex;
}

static pl_shader img_sh(struct pass_state *pass)
{
    if (pass->sh)
        return pass->sh;
    pass->sh = begin(pass);
    pl_shader_sample_direct(pass->sh, pl_sample_src( .tex = pass->tex ));
    pass->tex = NULL;
    return pass->sh;
}

// ---------------------------------------------------------------------------------------
// samplers (renderer.c:597-789)
// ---------------------------------------------------------------------------------------

enum sampler_type { SAMPLER_DIRECT, SAMPLER_NEAREST, SAMPLER_BICUBIC, SAMPLER_HERMITE,
                    SAMPLER_GAUSSIAN, SAMPLER_COMPLEX, SAMPLER_OVERSAMPLE };
enum sampler_dir { SAMPLER_NOOP, SAMPLER_UP, SAMPLER_DOWN };
enum sampler_usage { SAMPLER_MAIN, SAMPLER_PLANE, SAMPLER_LOWPASS };

struct sampler_info {
    const struct pl_filter_config *config;
    enum sampler_type type;
    enum sampler_dir dir, dir_sep[2];
};

static struct sampler_info sample_src_info(struct pass_state *pass,
                                           const struct pl_sample_src *src,
                                           enum sampler_usage usage)
{
    const struct pl_render_params *params = pass->params;
    struct sampler_info info = {0};
    const float rx = src->new_w / fabsf(pl_rect_w(src->rect));
    const float ry = src->new_h / fabsf(pl_rect_h(src->rect));
    if (rx < 1.0 - 1e-6)
        info.dir_sep[0] = SAMPLER_DOWN;
    else if (rx > 1.0 + 1e-6)
        info.dir_sep[0] = SAMPLER_UP;
    if (ry < 1.0 - 1e-6)
        info.dir_sep[1] = SAMPLER_DOWN;
    else if (ry > 1.0 + 1e-6)
        info.dir_sep[1] = SAMPLER_UP;
    if (params->correct_subpixel_offsets) {
        if (!info.dir_sep[0] && fabsf(src->rect.x0) > 1e-6f)
            info.dir_sep[0] = SAMPLER_UP;
        if (!info.dir_sep[1] && fabsf(src->rect.y0) > 1e-6f)
            info.dir_sep[1] = SAMPLER_UP;
    }
    info.dir = PLB_MAX(info.dir_sep[0], info.dir_sep[1]);
    switch (info.dir) {
    case SAMPLER_DOWN:
        info.config = usage == SAMPLER_LOWPASS ? &pl_filter_bicubic // renderer.c:630-631
                    : usage == SAMPLER_PLANE && params->plane_downscaler
                    ? params->plane_downscaler : params->downscaler;
        break;
    case SAMPLER_UP:
        info.config = usage == SAMPLER_PLANE && params->plane_upscaler
                    ? params->plane_upscaler : params->upscaler;
        break;
    case SAMPLER_NOOP:
        info.type = SAMPLER_NEAREST;
        return info;
    }
    if ((pass->rr->errors & PL_RENDER_ERR_SAMPLING) || !info.config) {
        info.type = SAMPLER_DIRECT;
    } else if (info.config->kernel == &pl_filter_function_oversample) {
        info.type = SAMPLER_OVERSAMPLE;
    } else {
        info.type = SAMPLER_COMPLEX;
        const bool can_fast = info.dir == SAMPLER_UP || params->skip_anti_aliasing;
        if (can_fast && !params->disable_builtin_scalers) {
            // the reference's built-in fast paths (renderer.c:657-676; every format of this
            // backend can be sampled linearly)
            if (pl_filter_config_eq(info.config, &pl_filter_bicubic))
                info.type = SAMPLER_BICUBIC;
            if (pl_filter_config_eq(info.config, &pl_filter_hermite))
                info.type = SAMPLER_HERMITE;
            if (pl_filter_config_eq(info.config, &pl_filter_gaussian))
                info.type = SAMPLER_GAUSSIAN;
            if (pl_filter_config_eq(info.config, &pl_filter_bilinear))
                info.type = SAMPLER_DIRECT;
            if (pl_filter_config_eq(info.config, &pl_filter_nearest))
                info.type = SAMPLER_NEAREST;
        }
    }
    return info;
}

// Records the right sampler into `sh`. For two-pass separable scaling the first pass is
// dispatched into an FBO here.
static bool dispatch_sampler(struct pass_state *pass, pl_shader sh, struct sampler_state *st,
                             enum sampler_usage usage, const struct pl_sample_src *src)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    const struct sampler_info info = sample_src_info(pass, src, usage);
    // by type only (renderer.c:702-724): an unscaled source is SAMPLER_NEAREST, not a bilinear
    // fetch at texel centres
    if (info.type == SAMPLER_DIRECT)
        return pl_shader_sample_direct(sh, src);
    if (info.type == SAMPLER_NEAREST)
        return pl_shader_sample_nearest(sh, src);
    if (info.type == SAMPLER_BICUBIC)
        return pl_shader_sample_bicubic(sh, src);
    if (info.type == SAMPLER_HERMITE)
        return pl_shader_sample_hermite(sh, src);
    if (info.type == SAMPLER_GAUSSIAN)
        return pl_shader_sample_gaussian(sh, src);
    if (info.type == SAMPLER_OVERSAMPLE)
        return pl_shader_sample_oversample(sh, src, info.config->kernel->params[0]);

    struct pl_sample_filter_params fp = {
        .filter = *info.config,
        .antiring = params->antiringing_strength,
        .no_widening = params->skip_anti_aliasing && usage != SAMPLER_LOWPASS,
        .lut = info.dir == SAMPLER_UP ? &st->up : &st->down,
    };
    bool ok;
    if (info.config->polar) {
        ok = pl_shader_sample_polar(sh, src, &fp);
    } else if (info.dir_sep[0] && info.dir_sep[1]) {
        // vertical pass into an FBO, horizontal pass into `sh` (renderer.c:742-768)
        struct pl_sample_src src1 = *src, src2 = *src;
        src1.new_w = src->tex->params.w;
        src1.rect.x0 = 0;
        src1.rect.x1 = src1.new_w;
        src2.rect.y0 = 0;
        src2.rect.y1 = src1.new_h;
        pl_shader tsh = begin(pass);
        ok = pl_shader_sample_ortho2(tsh, &src1, &fp);
        if (!ok) {
            pl_dispatch_abort(rr->dp, &tsh);
        } else {
            src2.tex = shader_to_tex(pass, &tsh, src1.new_w, src1.new_h, NULL, src->components);
            src2.scale = 1.0;
            ok = src2.tex && pl_shader_sample_ortho2(sh, &src2, &fp);
        }
    } else {
        ok = pl_shader_sample_ortho2(sh, src, &fp);
    }
    if (!ok) {
        PLB_ERR(rr->log, "Failed dispatching scaler..");
        rr->errors |= PL_RENDER_ERR_SAMPLING;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------
// pass_read_image (renderer.c:1553-1962)
// ---------------------------------------------------------------------------------------

struct plane_state {
    enum plane_type type;
    struct pl_plane plane;
    pl_tex tex;             // current texture (original, merged or FBO)
    pl_shader sh;           // pending single-plane shader (deband / grain)
    int w, h, comps;
    pl_fmt fmt;             // format whose sample_order applies
    struct pl_color_repr repr;
    pl_rect2df rect;
    int plane_w, plane_h;
};

// sh_subpass (src/shaders.c:676-686): a compute sub-shader can only be merged if its work group
// geometry is acceptable to the parent (sh_try_compute). The caller materialises the sub-shader
// into an FBO otherwise, exactly like pass_read_image does when sh_subpass returns NULL (:1868).
static bool compute_compatible(const pl_shader sh, bool sub_compute, int bw, int bh, bool flex)
{
    if (!sub_compute || !sh->is_compute || sh->wg_flexible || flex)
        return true;
    return sh->wg_w == bw && sh->wg_h == bh;
}

// Move the (single) plane of sub-shader `sub` into the plane list of `sh`, mapping its
// components into colour channels (sh_subpass + "color[map[c]] = tmp[c]", :1874-1891)
static bool merge_plane_shader(pl_shader sh, pl_shader sub, const int map[4], int comps,
                               float scale)
{
    if (sub->failed || sub->src != PLB_SRC_PLANES || sub->px.nplanes != 1 || sub->px.prog.n) {
        plb_shader_fail(sh, "internal: cannot merge this sub-shader");
        return false;
    }
    if (sh->px.nplanes >= 4) {
        plb_shader_fail(sh, "too many planes in one pass");
        return false;
    }
    if (sub->is_compute && !plb_shader_try_compute(sh, sub->wg_w, sub->wg_h, sub->wg_flexible)) {
        plb_shader_fail(sh, "internal: incompatible work group sizes (check compute_compatible first)");
        return false;
    }
    struct plb_plane_src *pl = &sh->px.planes[sh->px.nplanes++];
    *pl = sub->px.planes[0];
    pl->scale *= scale;
    for (int k = 0; k < 4; k++)
        pl->map[k] = k < comps ? map[k] : -1;
    // keep the LUT objects of the sub-shader alive
    for (int i = 0; i < sub->num_objs && sh->num_objs < (int) PLB_ARRAY_SIZE(sh->objs); i++) {
        sub->objs[i]->refcount++;
        sh->objs[sh->num_objs++] = sub->objs[i];
    }
    for (int i = 0; i < sub->num_steps; i++)
        plb_shader_describe(sh, sub->steps[i]);
    for (int i = 0; i < sub->num_texs; i++)
        plb_shader_use_tex(sh, sub->texs[i]);
    return true;
}

static bool needs_grain(const struct pass_state *pass)
{
    return pass->image.film_grain.type != PL_FILM_GRAIN_NONE;
}

static bool want_merge(struct pass_state *pass, const struct plane_state *st,
                       const struct plane_state *ref)
{
    const struct pl_render_params *params = pass->params;
    if (!(pass->rr->errors & PL_RENDER_ERR_DEBANDING) && params->deband_params)
        return true;
    struct pl_sample_src src = {
        .new_w = ref->w,
        .new_h = ref->h,
        .rect = { .x1 = st->w, .y1 = st->h },
    };
    if (sample_src_info(pass, &src, SAMPLER_PLANE).type == SAMPLER_COMPLEX)
        return true;
    return needs_grain(pass);
}

extern bool plb_plane_film_grain(pl_renderer rr, pl_shader sh, pl_shader_obj *state,
                                 const struct pl_frame *image, const struct pl_plane *plane,
                                 pl_tex tex, pl_tex luma_tex, int luma_comp,
                                 struct pl_color_repr *repr);

static bool pass_read_image(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    struct pl_frame *image = &pass->image;
    struct plane_state planes[4] = {0};
    struct plane_state *ref = &planes[pass->src_ref];

    for (int i = 0; i < image->num_planes; i++) {
        struct plane_state *st = &planes[i];
        st->type = detect_plane_type(&image->planes[i], &image->repr);
        st->plane = image->planes[i];
        st->tex = image->planes[i].texture;
        st->w = st->tex->params.w;
        st->h = st->tex->params.h;
        st->comps = image->planes[i].components;
        st->fmt = st->tex->params.format;
        st->repr = image->repr;
        if (st->plane.flipped) {
            PLB_ERR(rr->log, "pl_render_image: flipped planes are outside the CUDA hot path");
            return false;
        }
        if (image->repr.alpha == PL_ALPHA_NONE) {
            if (st->type == PLANE_ALPHA) {
                st->type = PLANE_INVALID;
                continue;
            }
            for (int j = 0; j < st->plane.components; j++) {
                if (st->plane.component_mapping[j] == PL_CHANNEL_A)
                    st->plane.component_mapping[j] = PL_CHANNEL_NONE;
            }
        }
    }
    // Deinterlace the planes first (renderer.c:1591-1612). On this backend every deinterlaced
    // plane goes through a texture of the plane's own format (the reference keeps it as a pending
    // shader and materialises it the same way as soon as debanding, scaling or grain needs a
    // texture; a plane merged directly would skip that rounding step).
    if (image->field != PL_FIELD_NONE && params->deinterlace_params &&
        !(rr->errors & PL_RENDER_ERR_DEINTERLACING)) {
        const struct pl_frame *refs[2] = { image->prev, image->next };
        for (int r = 0; r < 2; r++) {
            if (!refs[r])
                continue;
            bool ok = refs[r]->num_planes == image->num_planes;
            for (int i = 0; ok && i < image->num_planes; i++) {
                pl_tex a = refs[r]->planes[i].texture, b = image->planes[i].texture;
                ok = a && a->params.w == b->params.w && a->params.h == b->params.h &&
                     a->params.format == b->params.format;
            }
            if (!ok) {
                PLB_ERR(rr->log, "pl_render_image: deinterlacing reference frames must match the "
                        "image's planes (count, size, format)");
                return false;
            }
        }
        for (int i = 0; i < image->num_planes; i++) {
            struct plane_state *st = &planes[i];
            if (!st->type)
                continue;
            const struct pl_deinterlace_source src = {
                .cur.top = st->tex,
                .prev.top = image->prev ? image->prev->planes[i].texture : NULL,
                .next.top = image->next ? image->next->planes[i].texture : NULL,
                .field = image->field,
                .first_field = image->first_field,
                .component_mask = (1 << st->comps) - 1,
            };
            pl_tex out = get_fbo(pass, st->w, st->h, st->fmt, 0);
            pl_shader dsh = out ? begin(pass) : NULL;
            if (dsh)
                pl_shader_deinterlace(dsh, &src, params->deinterlace_params);
            if (!dsh || !pl_dispatch_finish(rr->dp, pl_dispatch_params( .shader = &dsh, .target = out ))) {
                PLB_ERR(rr->log, "Failed deinterlacing plane.. disabling!");
                rr->errors |= PL_RENDER_ERR_DEINTERLACING;
                if (dsh)
                    pl_dispatch_abort(rr->dp, &dsh);
                break;
            }
            st->tex = out;
        }
    }
    pl_tex ref_tex = ref->plane.texture;

    // --- A. merge compatible planes (chroma pairs) ---
    for (int i = 0; i < image->num_planes; i++) {
        struct plane_state *sti = &planes[i];
        if (!sti->type || !want_merge(pass, sti, ref))
            continue;
        for (int j = i + 1; j < image->num_planes; j++) {
            struct plane_state *stj = &planes[j];
            const bool merge = sti->type == stj->type && sti->w == stj->w && sti->h == stj->h &&
                               sti->plane.shift_x == stj->plane.shift_x &&
                               sti->plane.shift_y == stj->plane.shift_y;
            if (!merge || sti->fmt->type != stj->fmt->type)
                continue;
            const int num_comps = PLB_MIN(4, sti->comps + stj->comps);
            const int depth = PLB_MAX(sti->repr.bits.sample_depth, stj->repr.bits.sample_depth);
            pl_fmt fmt = sti->fmt->type == PL_FMT_UINT ? rr->fbofmt[num_comps]
                       : pl_find_fmt(rr->gpu, sti->fmt->type, num_comps, depth, 0,
                                     PL_FMT_CAP_SAMPLEABLE);
            if (!fmt)
                continue;
            // Two single-component planes of the target's component type: the merged texture
            // would hold exactly the same values, so alias them instead of copying
            if (sti->comps == 1 && stj->comps == 1 && num_comps == 2 && !sti->sh && !stj->sh &&
                sti->plane.component_mapping[0] != PL_CHANNEL_NONE &&
                stj->plane.component_mapping[0] != PL_CHANNEL_NONE && rr->num_views < 4) {
                pl_tex view = plb_tex_view2(rr->gpu, sti->tex, stj->tex, fmt);
                if (view) {
                    rr->views[rr->num_views++] = view;
                    sti->tex = view;
                    sti->fmt = fmt;
                    sti->comps = sti->plane.components = 2;
                    sti->plane.component_mapping[1] = stj->plane.component_mapping[0];
                    *stj = (struct plane_state) {0};
                    continue;
                }
            }
            // one pass that copies both planes into adjacent components
            pl_shader sh = begin(pass), sa = begin(pass), sb = begin(pass);
            pl_shader_sample_direct(sa, pl_sample_src( .tex = sti->tex ));
            pl_shader_sample_direct(sb, pl_sample_src( .tex = stj->tex ));
            int comp_map[4] = {-1, -1, -1, -1}, mapa[4] = {-1, -1, -1, -1}, mapb[4] = {-1, -1, -1, -1};
            int n = 0;
            for (int c = 0; c < sti->comps; c++) {
                if (sti->plane.component_mapping[c] == PL_CHANNEL_NONE)
                    continue;
                mapa[c] = n;
                comp_map[n++] = sti->plane.component_mapping[c];
            }
            for (int c = 0; c < stj->comps && n < 4; c++) {
                if (stj->plane.component_mapping[c] == PL_CHANNEL_NONE)
                    continue;
                mapb[c] = n;
                comp_map[n++] = stj->plane.component_mapping[c];
            }
            sh->src = PLB_SRC_PLANES;
            sh->px.coord_mode = PLB_COORD_EXACT;
            bool ok = merge_plane_shader(sh, sa, mapa, sti->comps, 1.0f) &&
                      merge_plane_shader(sh, sb, mapb, stj->comps, 1.0f);
            pl_dispatch_abort(rr->dp, &sa);
            pl_dispatch_abort(rr->dp, &sb);
            pl_tex merged = ok ? shader_to_tex(pass, &sh, sti->w, sti->h, fmt, n) : NULL;
            if (!merged) {
                pl_dispatch_abort(rr->dp, &sh);
                PLB_ERR(rr->log, "Failed dispatching plane merging shader");
                return false;
            }
            sti->tex = merged;
            sti->fmt = fmt;
            sti->comps = sti->plane.components = n;
            memcpy(sti->plane.component_mapping, comp_map, sizeof(comp_map));
            *stj = (struct plane_state) {0};
        }
    }

    // neutral values (:1715-1721)
    const int bits = image->repr.bits.sample_depth;
    const float out_scale = bits ? (1llu << bits) / ((1llu << bits) - 1.0f) : 1.0f;
    float neutral_luma = 0.0, neutral_chroma = 0.5f * out_scale;
    if (pl_color_levels_guess(&image->repr) == PL_COLOR_LEVELS_LIMITED)
        neutral_luma = 16 / 256.0f * out_scale;
    if (!pl_color_system_is_ycbcr_like(image->repr.sys))
        neutral_chroma = neutral_luma;

    // --- B. per-plane rects, debanding, grain ---
    for (int i = 0; i < image->num_planes; i++) {
        struct plane_state *st = &planes[i];
        if (!st->type)
            continue;
        const float rx = (float) st->plane.texture->params.w / ref_tex->params.w,
                    ry = (float) st->plane.texture->params.h / ref_tex->params.h;
        const float rrx = rx >= 1 ? roundf(rx) : 1.0 / roundf(1.0 / rx),
                    rry = ry >= 1 ? roundf(ry) : 1.0 / roundf(1.0 / ry);
        const float sx = st->plane.shift_x, sy = st->plane.shift_y;
        st->rect = (pl_rect2df) {
            (image->crop.x0 - sx) * rrx, (image->crop.y0 - sy) * rry,
            (image->crop.x1 - sx) * rrx, (image->crop.y1 - sy) * rry,
        };
        st->plane_w = ref_tex->params.w * rrx;
        st->plane_h = ref_tex->params.h * rry;

        float neutral[3] = {0.0};
        for (int c = 0, idx = 0; c < st->plane.components; c++) {
            switch (st->plane.component_mapping[c]) {
            case PL_CHANNEL_Y: neutral[idx++] = neutral_luma; break;
            case PL_CHANNEL_U:
            case PL_CHANNEL_V: neutral[idx++] = neutral_chroma; break;
            }
        }

        if (params->deband_params && !(rr->errors & PL_RENDER_ERR_DEBANDING)) {
            struct pl_color_repr repr = st->repr;
            struct pl_sample_src src = {
                .tex = st->tex,
                .components = st->comps,
                .scale = pl_color_repr_normalize(&repr),
            };
            // keep the grain amplitude independent of the source's nominal peak (:1336-1342)
            struct pl_deband_params dparams = *params->deband_params;
            dparams.grain /= image->color.hdr.max_luma / PL_COLOR_SDR_WHITE;
            memcpy(dparams.grain_neutral, neutral, sizeof(dparams.grain_neutral));
            st->sh = begin(pass);
            pl_shader_deband(st->sh, &src, &dparams);
            if (pl_shader_is_failed(st->sh)) {
                pl_dispatch_abort(rr->dp, &st->sh);
                rr->errors |= PL_RENDER_ERR_DEBANDING;
            } else {
                st->repr = repr;
            }
        }

        if (needs_grain(pass)) {
            // grain needs the (possibly debanded) plane as a texture
            if (st->sh) {
                st->tex = shader_to_tex(pass, &st->sh, st->w, st->h, NULL, st->comps);
                st->sh = NULL;
                st->fmt = st->tex ? st->tex->params.format : st->fmt;
                if (!st->tex)
                    return false;
            }
            int luma_comp = 0;
            for (int c = 0; c < ref->plane.components; c++) {
                if (ref->plane.component_mapping[c] == PL_CHANNEL_Y)
                    luma_comp = c;
            }
            pl_shader gsh = begin(pass);
            // renderer.c:1366-1376: the grain always sees the ORIGINAL colour depth; if the
            // plane was already processed (normalised), only sample depth / shift follow it
            struct pl_color_repr grepr = image->repr;
            const bool is_orig_repr = pl_color_repr_equal(&st->repr, &image->repr);
            if (!is_orig_repr) {
                grepr.bits.sample_depth = st->repr.bits.sample_depth;
                grepr.bits.bit_shift = grepr.bits.sample_depth - grepr.bits.color_depth;
            }
            if (plb_plane_film_grain(rr, gsh, &rr->grain_state[i], image, &st->plane, st->tex,
                                     ref_tex, luma_comp, &grepr)) {
                st->sh = gsh;
                if (is_orig_repr)
                    st->repr = grepr;
            } else {
                const bool failed = pl_shader_is_failed(gsh);
                pl_dispatch_abort(rr->dp, &gsh);
                if (failed) {
                    rr->errors |= PL_RENDER_ERR_FILM_GRAIN;
                    PLB_ERR(rr->log, "Failed applying film grain..");
                    return false;
                }
            }
        }
    }

    // --- C. align all planes onto the reference grid and merge them ---
    pl_shader sh = begin(pass);
    sh->px.init[0] = neutral_luma;
    sh->px.init[1] = sh->px.init[2] = neutral_chroma;
    sh->px.init[3] = 1.0f;
    sh->px.coord_mode = plb_gpu_priv(rr->gpu)->coord_mode;

    const pl_rect2df ref_rc = ref->rect;
    pl_rect2d ref_rounded;
    ref_rounded.x0 = truncf(ref_rc.x0);
    ref_rounded.y0 = truncf(ref_rc.y0);
    ref_rounded.x1 = ref_rounded.x0 + roundf(pl_rect_w(ref_rc));
    ref_rounded.y1 = ref_rounded.y0 + roundf(pl_rect_h(ref_rc));
    const float off_x = ref_rc.x0 - ref_rounded.x0, off_y = ref_rc.y0 - ref_rounded.y0,
                stretch_x = pl_rect_w(ref_rounded) / pl_rect_w(ref_rc),
                stretch_y = pl_rect_h(ref_rounded) / pl_rect_h(ref_rc);
    const int out_w = pl_rect_w(ref_rounded), out_h = pl_rect_h(ref_rounded);

    for (int i = 0; i < image->num_planes; i++) {
        struct plane_state *st = &planes[i];
        if (!st->type)
            continue;
        const float scale_x = pl_rect_w(st->rect) / pl_rect_w(ref_rc),
                    scale_y = pl_rect_h(st->rect) / pl_rect_h(ref_rc),
                    base_x = st->rect.x0 - scale_x * off_x,
                    base_y = st->rect.y0 - scale_y * off_y;
        struct pl_sample_src src = {
            .components = st->plane.components,
            .address_mode = st->plane.address_mode,
            .scale = pl_color_repr_normalize(&st->repr),
            .new_w = out_w,
            .new_h = out_h,
            .rect = {
                base_x, base_y,
                base_x + stretch_x * pl_rect_w(st->rect),
                base_y + stretch_y * pl_rect_h(st->rect),
            },
        };
        int map[4] = {-1, -1, -1, -1};
        for (int c = 0; c < src.components; c++)
            map[st->fmt->sample_order[c]] = st->plane.component_mapping[c];

        const pl_rect2d unscaled = { .x1 = out_w, .y1 = out_h };
        const bool aligned = st->w == out_w && st->h == out_h && pl_rect2d_eq(src.rect, unscaled);
        bool ok;
        if (st->sh && aligned && !compute_compatible(sh, st->sh->is_compute, st->sh->wg_w,
                                                      st->sh->wg_h, st->sh->wg_flexible)) {
            // sh_subpass refuses (work group clash): through an FBO, then a plain fetch
            st->tex = shader_to_tex(pass, &st->sh, st->w, st->h, NULL, st->comps);
            st->sh = NULL;
            if (!st->tex)
                goto error;
            pl_shader psh = begin(pass);
            ok = pl_shader_sample_direct(psh, pl_sample_src( .tex = st->tex )) &&
                 merge_plane_shader(sh, psh, map, src.components, 1.0f);
            if (sh->src == PLB_SRC_NONE)
                sh->src = PLB_SRC_PLANES;
            pl_dispatch_abort(rr->dp, &psh);
        } else if (st->sh && aligned) {
            // pending 1:1 plane shader (deband / grain): merge as is
            ok = merge_plane_shader(sh, st->sh, map, src.components, src.scale);
            if (sh->src == PLB_SRC_NONE)
                sh->src = PLB_SRC_PLANES;
            pl_dispatch_abort(rr->dp, &st->sh);
        } else {
            if (st->sh) {
                st->tex = shader_to_tex(pass, &st->sh, st->w, st->h, NULL, st->comps);
                st->sh = NULL;
                if (!st->tex)
                    goto error;
            }
            src.tex = st->tex;
            const struct sampler_info info = sample_src_info(pass, &src, SAMPLER_PLANE);
            // a polar sampler is a rigid 32x32 compute shader (sampling.c:668-699): it cannot
            // join a pass that already has another rigid geometry (e.g. H.274 grain, 8x8)
            const bool clash = info.type == SAMPLER_COMPLEX && info.config->polar &&
                               !compute_compatible(sh, true, 32, 32, false);
            if (clash) {
                pl_shader psh = begin(pass);
                ok = dispatch_sampler(pass, psh, &rr->samplers_src[i], SAMPLER_PLANE, &src);
                pl_tex t = ok ? shader_to_tex(pass, &psh, out_w, out_h, NULL, st->comps) : NULL;
                psh = begin(pass);
                ok = t && pl_shader_sample_direct(psh, pl_sample_src( .tex = t )) &&
                     merge_plane_shader(sh, psh, map, src.components, 1.0f);
                if (sh->src == PLB_SRC_NONE)
                    sh->src = PLB_SRC_PLANES;
                pl_dispatch_abort(rr->dp, &psh);
            } else if (info.type == SAMPLER_COMPLEX && sh->src != PLB_SRC_POLAR && sh->src != PLB_SRC_ORTHO) {
                // this plane's scaler becomes the primary sampler of the merged pass
                const enum plb_src_kind had = sh->src;
                sh->src = PLB_SRC_NONE;
                ok = dispatch_sampler(pass, sh, &rr->samplers_src[i], SAMPLER_PLANE, &src);
                if (ok && (sh->src == PLB_SRC_POLAR || sh->src == PLB_SRC_ORTHO)) {
                    int8_t *smap = sh->src == PLB_SRC_POLAR ? sh->polar.src_map : sh->ortho.src_map;
                    for (int k = 0; k < 4; k++)
                        smap[k] = k < src.components ? map[k] : -1;
                } else if (ok) {
                    // fell back to direct sampling: turn it into a plane entry
                    struct plb_plane_src tmp = sh->px.planes[0];
                    sh->px.nplanes--;
                    sh->px.planes[sh->px.nplanes] = tmp;
                    for (int k = 0; k < 4; k++)
                        sh->px.planes[sh->px.nplanes].map[k] = k < src.components ? map[k] : -1;
                    sh->px.nplanes++;
                    sh->src = had ? had : PLB_SRC_PLANES;
                }
                sh->out_w = sh->out_h = 0; // the merged pass is sized by the reference plane
            } else {
                pl_shader psh = begin(pass);
                if (info.type == SAMPLER_COMPLEX) {
                    // a second scaled plane: materialise at full precision, then merge 1:1
                    ok = dispatch_sampler(pass, psh, &rr->samplers_src[i], SAMPLER_PLANE, &src);
                    pl_fmt f32 = pl_find_fmt(rr->gpu, PL_FMT_FLOAT, st->comps == 3 ? 4 : st->comps,
                                             32, 0, PL_FMT_CAP_SAMPLEABLE);
                    pl_tex t = ok ? shader_to_tex(pass, &psh, out_w, out_h, f32, st->comps) : NULL;
                    psh = begin(pass);
                    ok = t && pl_shader_sample_direct(psh, pl_sample_src( .tex = t ));
                    src.scale = 1.0;
                } else {
                    // texture-unit scalers and built-in fast paths; the plane's scale factor
                    // is applied by the merge ("tmp = vec4(scale) * sub()")
                    struct pl_sample_src plain = {
                        .tex = src.tex, .rect = src.rect, .new_w = out_w, .new_h = out_h };
                    ok = dispatch_sampler(pass, psh, &rr->samplers_src[i], SAMPLER_PLANE, &plain);
                }
                ok = ok && merge_plane_shader(sh, psh, map, src.components, src.scale);
                if (sh->src == PLB_SRC_NONE)
                    sh->src = PLB_SRC_PLANES;
                pl_dispatch_abort(rr->dp, &psh);
            }
        }
        if (!ok)
            goto error;
    }

    pass->sh = sh;
    pass->tex = NULL;
    pass->w = out_w;
    pass->h = out_h;
    pass->repr = ref->repr;
    pass->color = image->color;
    pass->comps = ref->repr.alpha == PL_ALPHA_NONE ? 3 : 4;
    pass->rect = (pl_rect2df) { off_x, off_y, off_x + pl_rect_w(ref_rc), off_y + pl_rect_h(ref_rc) };
    pass->ref_rect = pass->rect;

    // LUT logic and colour decoding (renderer.c:1919-1946)
    const enum pl_lut_type lut_type = guess_frame_lut_type(image, false);
    bool needs_conversion = true;
    if (lut_type == PL_LUT_NATIVE || lut_type == PL_LUT_CONVERSION) {
        // fix the bit depth normalisation before applying the LUT
        const float scale = pl_color_repr_normalize(&pass->repr);
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_SCALE);
        if (!op)
            goto error;
        op->u.scale.s = scale;
        op->u.scale.with_alpha = 1;
        pl_shader_custom_lut(sh, image->lut, &rr->lut_state[LUT_IMAGE]);
        if (lut_type == PL_LUT_CONVERSION) {
            pass->repr.sys = PL_COLOR_SYSTEM_RGB;
            pass->repr.levels = PL_COLOR_LEVELS_FULL;
            needs_conversion = false;
        }
    }
    if (needs_conversion) {
        if (pass->repr.sys == PL_COLOR_SYSTEM_XYZ) {
            pl_shader_linearize(sh, &pass->color);
            pass->color.transfer = PL_COLOR_TRC_LINEAR;
        }
        pl_shader_decode_color(sh, &pass->repr, params->color_adjustment);
    }
    if (lut_type == PL_LUT_NORMALIZED)
        pl_shader_custom_lut(sh, image->lut, &rr->lut_state[LUT_IMAGE]);
    pl_shader_set_alpha(sh, &pass->repr, PL_ALPHA_PREMULTIPLIED);
    return !pl_shader_is_failed(sh);

error:
    pl_dispatch_abort(rr->dp, &sh);
    for (int i = 0; i < 4; i++)
        pl_dispatch_abort(rr->dp, &planes[i].sh);
    return false;
}

// ---------------------------------------------------------------------------------------
// hdr_update_peak (renderer.c:1183-1250)
// ---------------------------------------------------------------------------------------

static void hdr_update_peak(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    if (!params->peak_detect_params || !pl_color_space_is_hdr(&pass->image.color))
        goto cleanup;
    if (rr->errors & PL_RENDER_ERR_PEAK_DETECT)
        goto cleanup;
    float max_peak = pl_color_transfer_nominal_peak(pass->image.color.transfer) * PL_COLOR_SDR_WHITE;
    if (pass->image.color.transfer == PL_COLOR_TRC_HLG)
        max_peak = pass->color.hdr.max_luma;
    if (max_peak <= pass->target.color.hdr.max_luma + 1e-6)
        goto cleanup; // nothing to adapt
    if (pass->color.hdr.avg_pq_y)
        goto cleanup; // dynamic metadata already present
    if (params->lut && params->lut_type == PL_LUT_CONVERSION)
        goto cleanup; // LUT handles tone mapping
    const struct pl_color_map_params *cpars = params->color_map_params;
    const enum pl_hdr_metadata_type metadata = cpars ? cpars->metadata : PL_HDR_METADATA_ANY;
    if (metadata && metadata != PL_HDR_METADATA_CIE_Y)
        goto cleanup;
    const bool uses_ootf = cpars && cpars->tone_mapping_function == &pl_tone_map_st2094_40;
    if (uses_ootf && pass->color.hdr.ootf.num_anchors)
        goto cleanup;
    if (rr->have_peak_override) {
        // frame-parallel mode: the caller supplies the smoothed state (renderer.h extension);
        // keep the FBO round trip of the sequential same-frame flow so that the pixels are
        // identical to a single-GPU stream
        // (continuing from pl_renderer_measure_peak's image, that round trip has happened)
        pass->need_peak_fbo = !params->peak_detect_params->allow_delayed && !pass->from_measured;
        goto cleanup_keep;
    }
    if (!pl_shader_detect_peak(img_sh(pass), pass->color, &rr->tone_map_state,
                               params->peak_detect_params)) {
        PLB_WARN(rr->log, "Failed creating HDR peak detection shader.. disabling");
        rr->errors |= PL_RENDER_ERR_PEAK_DETECT;
        goto cleanup;
    }
    pass->need_peak_fbo = !params->peak_detect_params->allow_delayed;
    return;

cleanup:
    pl_reset_detected_peak(rr->tone_map_state);
cleanup_keep:
    return;
}

// ---------------------------------------------------------------------------------------
// pass_scale_main (renderer.c:1964-2087)
// ---------------------------------------------------------------------------------------

static bool pass_scale_main(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    const pl_rect2df new_rect = {
        .x1 = abs(pl_rect_w(pass->dst_rect)),
        .y1 = abs(pl_rect_h(pass->dst_rect)),
    };
    struct pl_sample_src src = {
        .components = pass->comps,
        .new_w = pl_rect_w(new_rect),
        .new_h = pl_rect_h(new_rect),
        .rect = pass->rect,
    };
    const struct sampler_info info = sample_src_info(pass, &src, SAMPLER_MAIN);
    bool use_sigmoid = info.dir == SAMPLER_UP && params->sigmoid_params;
    bool use_linear = info.dir == SAMPLER_DOWN;
    if (info.dir == SAMPLER_UP)
        hdr_update_peak(pass);

    if (info.dir == SAMPLER_NOOP) {
        // 1:1: nothing to do (src and dst sizes agree by construction)
        goto done;
    }
    if (info.type == SAMPLER_DIRECT) {
        // "free" bilinear resize: needs the image as a texture anyway on this backend
        use_sigmoid = use_linear = false;
    }

    pl_fmt fbofmt = rr->fbofmt[pass->comps];
    if (params->disable_linear_scaling || fbofmt->component_depth[0] < 16)
        use_sigmoid = use_linear = false;
    if (pl_color_space_is_hdr(&pass->color)) {
        use_sigmoid = false;
        if (fbofmt->type != PL_FMT_FLOAT)
            use_linear = false;
    }
    if (use_linear || use_sigmoid) {
        pl_shader_linearize(img_sh(pass), &pass->color);
        pass->color.transfer = PL_COLOR_TRC_LINEAR;
    }
    if (use_sigmoid)
        pl_shader_sigmoidize(img_sh(pass), params->sigmoid_params);

    src.tex = img_tex(pass);
    if (!src.tex)
        return false;
    pass->need_peak_fbo = false;

    pl_shader sh = begin(pass);
    if (!dispatch_sampler(pass, sh, &rr->sampler_main, SAMPLER_MAIN, &src)) {
        pl_dispatch_abort(rr->dp, &sh);
        return false;
    }
    pass->tex = NULL;
    pass->sh = sh;
    pass->w = src.new_w;
    pass->h = src.new_h;
    pass->rect = new_rect;
    if (use_sigmoid)
        pl_shader_unsigmoidize(sh, params->sigmoid_params);

done:
    if (info.dir != SAMPLER_UP)
        hdr_update_peak(pass);
    return true;
}

// ---------------------------------------------------------------------------------------
// pass_convert_colors (renderer.c:2157-2279)
// ---------------------------------------------------------------------------------------

// get_feature_map (renderer.c:2089-2154): the low-pass "feature map" of contrast recovery — the
// image's PQ intensity (pl_shader_extract_features) at full size into a one-component FBO, then
// a bicubic down-scaling (widened kernel, mirrored borders) by contrast_smoothness.
static pl_tex get_feature_map(struct pass_state *pass)
{
    const struct pl_render_params *params = pass->params;
    pl_renderer rr = pass->rr;
    const struct pl_color_map_params *cparams = PLB_DEF(params->color_map_params, &pl_color_map_default_params);
    if (!cparams->contrast_recovery || cparams->contrast_smoothness <= 1)
        return NULL;
    if (!pl_color_space_is_hdr(&pass->color))
        return NULL;
    if (rr->errors & (PL_RENDER_ERR_SAMPLING | PL_RENDER_ERR_CONTRAST_RECOVERY))
        return NULL;
    if (pass->color.hdr.max_luma <= pass->target.color.hdr.max_luma + 1e-6)
        return NULL; // no adaptation needed
    if (params->lut && params->lut_type == PL_LUT_CONVERSION)
        return NULL; // LUT handles tone mapping
    pl_tex img = img_tex(pass);
    if (!img)
        return NULL;

    const float ratio = cparams->contrast_smoothness;
    const int cr_w = ceilf(abs(pl_rect_w(pass->dst_rect)) / ratio);
    const int cr_h = ceilf(abs(pl_rect_h(pass->dst_rect)) / ratio);
    pl_tex inter_tex = get_fbo(pass, pass->w, pass->h, NULL, 1);
    pl_tex out_tex = get_fbo(pass, cr_w, cr_h, NULL, 1);
    if (!inter_tex || !out_tex)
        goto error;

    pl_shader sh = begin(pass);
    pl_shader_sample_direct(sh, pl_sample_src( .tex = img ));
    pl_shader_extract_features(sh, pass->color);
    if (!pl_dispatch_finish(rr->dp, pl_dispatch_params( .shader = &sh, .target = inter_tex )))
        goto error;

    const struct pl_sample_src src = {
        .tex = inter_tex,
        .rect = pass->rect,
        .address_mode = PL_TEX_ADDRESS_MIRROR,
        .components = 1,
        .new_w = cr_w,
        .new_h = cr_h,
    };
    sh = begin(pass);
    if (!dispatch_sampler(pass, sh, &rr->sampler_contrast, SAMPLER_LOWPASS, &src)) {
        pl_dispatch_abort(rr->dp, &sh);
        goto error;
    }
    if (!pl_dispatch_finish(rr->dp, pl_dispatch_params( .shader = &sh, .target = out_tex )))
        goto error;
    return out_tex;

error:
    PLB_ERR(rr->log, "Failed extracting luma for contrast recovery, disabling");
    rr->errors |= PL_RENDER_ERR_CONTRAST_RECOVERY;
    return NULL;
}

static bool pass_convert_colors(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    const struct pl_frame *image = &pass->image, *target = &pass->target;
    pl_shader sh = img_sh(pass);
    bool prelinearized = false;
    if (pass->color.transfer == PL_COLOR_TRC_LINEAR) {
        if (pass->repr.alpha == PL_ALPHA_PREMULTIPLIED) {
            pass->color.transfer = image->color.transfer;
            pl_shader_delinearize(sh, &pass->color);
        } else {
            prelinearized = true;
        }
    } else if (pass->color.transfer != image->color.transfer) {
        if (image->color.transfer == PL_COLOR_TRC_LINEAR) {
            pl_shader_linearize(sh, &pass->color);
            pass->color.transfer = PL_COLOR_TRC_LINEAR;
        }
    }
    pl_shader_set_alpha(sh, &pass->repr, PL_ALPHA_INDEPENDENT);

    // the main LUT (renderer.c:2198-2247): image colour space -> LUT input, the LUT, LUT output ->
    // image colour space again (a CONVERSION LUT replaces the conversion to the target instead)
    bool need_conversion = true;
    if (params->lut) {
        struct pl_color_space lut_in = params->lut->color_in, lut_out = params->lut->color_out;
        switch (params->lut_type) {
        case PL_LUT_UNKNOWN:
        case PL_LUT_NATIVE:
            pl_color_space_merge(&lut_in, &image->color);
            pl_color_space_merge(&lut_out, &image->color);
            break;
        case PL_LUT_CONVERSION:
            pl_color_space_merge(&lut_in, &image->color);
            need_conversion = false; // conversion LUT the highest priority
            break;
        case PL_LUT_NORMALIZED:
            if (!prelinearized) {
                pl_shader_linearize(sh, &pass->color); // PL_LUT_NORMALIZED wants linear input data
                pass->color.transfer = PL_COLOR_TRC_LINEAR;
                prelinearized = true;
            }
            pl_color_space_merge(&lut_in, &pass->color);
            pl_color_space_merge(&lut_out, &pass->color);
            break;
        }
        pl_shader_color_map_ex(sh, params->color_map_params, pl_color_map_args(
            .src = image->color,
            .dst = lut_in,
            .prelinearized = prelinearized,
        ));
        struct plb_color_op *op;
        if (params->lut_type == PL_LUT_NORMALIZED) {
            if (!(op = plb_shader_push_op(sh, PLB_OP_SCALE)))
                return false;
            op->u.scale.s = 1.0f / pl_color_transfer_nominal_peak(lut_in.transfer);
        }
        pl_shader_custom_lut(sh, params->lut, &rr->lut_state[LUT_PARAMS]);
        if (params->lut_type == PL_LUT_NORMALIZED) {
            if (!(op = plb_shader_push_op(sh, PLB_OP_SCALE)))
                return false;
            op->u.scale.s = pl_color_transfer_nominal_peak(lut_out.transfer);
        }
        if (params->lut_type != PL_LUT_CONVERSION) {
            pl_shader_color_map_ex(sh, params->color_map_params, pl_color_map_args(
                .src = lut_out,
                .dst = pass->color,
            ));
        }
    }

    if (need_conversion) {
        if (pass->need_peak_fbo && !img_tex(pass))
            return false;
        pl_tex feature_map = get_feature_map(pass); // dispatches the pending image shader
        sh = img_sh(pass);

        struct pl_color_space src_csp = image->color;
        if (rr->have_peak_override) {
            // the reference would read these from its own smoothed state (colorspace.c:1620-1621)
            src_csp.hdr.max_pq_y = rr->peak_override.max_pq;
            src_csp.hdr.avg_pq_y = rr->peak_override.avg_pq;
        }
        pl_shader_color_map_ex(sh, params->color_map_params, pl_color_map_args(
            .src = src_csp,
            .dst = target->color,
            .prelinearized = prelinearized,
            .state = &rr->tone_map_state,
            .feature_map = feature_map,
        ));
    }

    const enum pl_lut_type tlut = guess_frame_lut_type(target, true);
    if (tlut == PL_LUT_NORMALIZED || tlut == PL_LUT_CONVERSION)
        pl_shader_custom_lut(sh, target->lut, &rr->lut_state[LUT_TARGET]);
    pass->color = target->color;
    return !pl_shader_is_failed(sh);
}

// ---------------------------------------------------------------------------------------
// pass_output_target (renderer.c:2586-2964)
// ---------------------------------------------------------------------------------------

extern bool plb_record_swizzle(pl_shader sh, int comps, const int comp_map[4], bool force_alpha);
extern bool plb_pass_error_diffusion(pl_renderer rr, pl_dispatch dp, pl_shader *sh, int new_depth,
                                     int comps, int out_w, int out_h, pl_fmt fbofmt,
                                     const struct pl_error_diffusion_kernel *kernel,
                                     pl_tex (*get_fbo)(void *, int, int, pl_fmt, int), void *priv);

static pl_tex get_fbo_cb(void *priv, int w, int h, pl_fmt fmt, int comps)
{
    return get_fbo(priv, w, h, fmt, comps);
}

static bool pass_output_target(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    const struct pl_render_params *params = pass->params;
    const struct pl_frame *target = &pass->target;
    const struct pl_plane *ref = &target->planes[pass->dst_ref];
    if (pl_frame_is_cropped(target) && params->border != PL_CLEAR_SKIP) {
        // clear the target to the background colour first (clear_target :2516-2584); only
        // black / the configured solid colour is supported, packed targets only
        if (target->num_planes != 1) {
            PLB_ERR(rr->log, "pl_render_image: cropped planar targets need border clearing in the "
                    "target's colour system, which is outside the CUDA hot path (use PL_CLEAR_SKIP)");
            return false;
        }
        float bg[4] = { params->background_color[0], params->background_color[1],
                        params->background_color[2], 1.0 - params->background_transparency };
        pl_tex_clear(rr->gpu, target->planes[0].texture, bg);
    }

    pl_shader sh = img_sh(pass);

    // alpha handling for opaque sources / targets (:2713-2761): the hot path has no alpha
    if (pass->comps == 4) {
        PLB_ERR(rr->log, "pl_render_image: alpha blending is outside the CUDA hot path");
        return false;
    }

    struct pl_color_repr repr = target->repr;
    const float scale = pl_color_repr_normalize(&repr);
    repr.alpha = PL_ALPHA_NONE; // img->comps < 4
    const enum pl_lut_type lut_type = guess_frame_lut_type(target, true);
    if (lut_type != PL_LUT_CONVERSION) {
        pl_shader_encode_color(sh, &repr);
        if (repr.sys == PL_COLOR_SYSTEM_XYZ) {
            pass->color.transfer = PL_COLOR_TRC_ST428;
            pl_shader_delinearize(sh, &pass->color);
        }
    }
    if (lut_type == PL_LUT_NATIVE)
        pl_shader_custom_lut(sh, target->lut, &rr->lut_state[LUT_TARGET]);
    sh = NULL;

    // one pass per target plane (:2796-2961)
    for (int p = 0; p < target->num_planes; p++) {
        const struct pl_plane *plane = &target->planes[p];
        const float rx = (float) plane->texture->params.w / ref->texture->params.w,
                    ry = (float) plane->texture->params.h / ref->texture->params.h;
        // only integer scaling ratios: fractionally subsampled planes are rounded up to the
        // next integer size and over-rendered
        const float rrx = rx >= 1 ? roundf(rx) : 1.0 / roundf(1.0 / rx),
                    rry = ry >= 1 ? roundf(ry) : 1.0 / roundf(1.0 / ry);
        const float sx = plane->shift_x, sy = plane->shift_y;
        pl_rect2df plane_rectf = {
            .x0 = (pass->dst_rect.x0 - sx) * rrx,
            .y0 = (pass->dst_rect.y0 - sy) * rry,
            .x1 = (pass->dst_rect.x1 - sx) * rrx,
            .y1 = (pass->dst_rect.y1 - sy) * rry,
        };
        const int rx0 = floorf(plane_rectf.x0), ry0 = floorf(plane_rectf.y0),
                  rx1 = ceilf(plane_rectf.x1), ry1 = ceilf(plane_rectf.y1);
        if (plane->flipped) {
            PLB_ERR(rr->log, "pl_render_image: flipped target planes are outside the CUDA hot path");
            return false;
        }

        if (target->num_planes > 1) {
            // planar output: every plane samples the encoded image from an intermediate FBO
            pl_tex tex = img_tex(pass);
            if (!tex) {
                PLB_ERR(rr->log, "Output requires multiple planes, but FBOs are unavailable. "
                        "This combination is unsupported.");
                return false;
            }
            uint8_t mask = 0;
            for (int c = 0; c < plane->components; c++) {
                if (plane->component_mapping[c] >= 0)
                    mask |= 1 << plane->component_mapping[c];
            }
            const struct pl_sample_src src = {
                .tex = tex,
                .new_w = rx1 - rx0,
                .new_h = ry1 - ry0,
                .rect = {
                    .x0 = (rx0 - plane_rectf.x0) / rrx,
                    .x1 = (rx1 - plane_rectf.x0) / rrx,
                    .y0 = (ry0 - plane_rectf.y0) / rry,
                    .y1 = (ry1 - plane_rectf.y0) / rry,
                },
                .component_mask = mask,
            };
            sh = begin(pass);
            if (!dispatch_sampler(pass, sh, &rr->samplers_dst[p], SAMPLER_PLANE, &src)) {
                pl_dispatch_abort(rr->dp, &sh);
                return false;
            }
        } else {
            sh = img_sh(pass);
            pass->sh = NULL;
        }

        const int depth = target->repr.bits.color_depth;
        int applied_dither = 0;
        if (depth && (depth < 16 || params->force_dither)) {
            if (params->error_diffusion && !(rr->errors & PL_RENDER_ERR_ERROR_DIFFUSION)) {
                if (plb_pass_error_diffusion(rr, rr->dp, &sh, depth, plane->components, rx1 - rx0,
                                             ry1 - ry0, rr->fbofmt[4], params->error_diffusion,
                                             get_fbo_cb, pass)) {
                    applied_dither = depth;
                } else {
                    rr->errors |= PL_RENDER_ERR_ERROR_DIFFUSION;
                    if (!sh)
                        return false;
                }
            }
            if (!applied_dither && params->dither_params) {
                struct pl_dither_params dparams = *params->dither_params;
                if (!params->disable_dither_gamma_correction)
                    dparams.transfer = target->color.transfer;
                pl_shader_dither(sh, depth, &rr->dither_state, &dparams);
                applied_dither = depth;
            }
        }
        if (applied_dither != rr->prev_dither) {
            if (applied_dither)
                PLB_INFO(rr->log, "Dithering to %d bit depth", applied_dither);
            else
                PLB_INFO(rr->log, "Dithering disabled");
            rr->prev_dither = applied_dither;
        }

        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_SCALE);
        if (!op) {
            pl_dispatch_abort(rr->dp, &sh);
            return false;
        }
        op->u.scale.s = (float) (1.0 / scale);
        op->u.scale.with_alpha = 1;
        if (!plb_record_swizzle(sh, plane->components, plane->component_mapping, false)) {
            pl_dispatch_abort(rr->dp, &sh);
            return false;
        }
        // The rounded rect of a subsampled plane may overhang its texture by a pixel (the
        // reference over-renders and lets the rasteriser clip, dispatch.c:1290-1296): render
        // such planes into a scratch texture of the plane's format and copy the visible part.
        const int tw = plane->texture->params.w, th = plane->texture->params.h;
        const bool overhang = rx0 < 0 || ry0 < 0 || rx1 > tw || ry1 > th;
        pl_tex out = plane->texture;
        pl_rect2d rc = { rx0, ry0, rx1, ry1 };
        if (overhang) {
            out = get_fbo(pass, rx1 - rx0, ry1 - ry0, plane->texture->params.format, 0);
            if (!out) {
                pl_dispatch_abort(rr->dp, &sh);
                return false;
            }
            rc = (pl_rect2d) { 0, 0, rx1 - rx0, ry1 - ry0 };
        }
        if (!pl_dispatch_finish(rr->dp, pl_dispatch_params( .shader = &sh, .target = out, .rect = rc )))
            return false;
        if (overhang) {
            extern bool plb_tex_copy_region(pl_gpu gpu, pl_tex dst, int dx, int dy, pl_tex src,
                                            int sx, int sy, int w, int h);
            const int cx0 = PLB_MAX(rx0, 0), cy0 = PLB_MAX(ry0, 0),
                      cx1 = PLB_MIN(rx1, tw), cy1 = PLB_MIN(ry1, th);
            if (cx1 > cx0 && cy1 > cy0 &&
                !plb_tex_copy_region(rr->gpu, plane->texture, cx0, cy0, out, cx0 - rx0, cy0 - ry0,
                                     cx1 - cx0, cy1 - cy0))
                return false;
        }
    }
    pass->sh = NULL;
    pass->tex = NULL;
    return true;
}

// ---------------------------------------------------------------------------------------
// entry points
// ---------------------------------------------------------------------------------------

static bool check_params(pl_renderer rr, const struct pl_render_params *p,
                         const struct pl_frame *image, const struct pl_frame *target)
{
#define REJECT(cond, what) \
    if (cond) { \
        PLB_ERR(rr->log, "pl_render_image: %s is outside the CUDA hot path (SURVEY.md §2)", what); \
        return false; \
    }
    REJECT(p->num_hooks, "user hooks");
    REJECT(image->icc || target->icc || image->profile.len || target->profile.len, "ICC profiles");
    REJECT(image->num_overlays || target->num_overlays, "overlays");
    REJECT(p->cone_params, "cone distortion");
    REJECT(p->blend_params, "blending");
    REJECT(p->distort_params, "distortion");
    REJECT(p->corner_rounding > 0, "corner rounding");
    REJECT(p->border == PL_CLEAR_BLUR, "blurred borders");
#undef REJECT
    if (!image->num_planes || image->num_planes > PL_MAX_PLANES ||
        !target->num_planes || target->num_planes > PL_MAX_PLANES) {
        PLB_ERR(rr->log, "pl_render_image: invalid number of planes");
        return false;
    }
    for (int i = 0; i < image->num_planes; i++) {
        const struct pl_plane *pl = &image->planes[i];
        if (!pl->texture || !pl->texture->params.sampleable || pl->components <= 0 ||
            pl->components > 4) {
            PLB_ERR(rr->log, "pl_render_image: invalid source plane %d", i);
            return false;
        }
    }
    for (int i = 0; i < target->num_planes; i++) {
        const struct pl_plane *pl = &target->planes[i];
        if (!pl->texture || !pl->texture->params.renderable || pl->components <= 0 ||
            pl->components > 4) {
            PLB_ERR(rr->log, "pl_render_image: invalid target plane %d", i);
            return false;
        }
    }
    if (!pl_rect_w(image->crop) != !pl_rect_h(image->crop) ||
        !pl_rect_w(target->crop) != !pl_rect_h(target->crop)) {
        PLB_ERR(rr->log, "pl_render_image: inconsistent crop");
        return false;
    }
    return true;
}

static bool pass_begin(struct pass_state *pass)
{
    pl_renderer rr = pass->rr;
    if (!check_params(rr, pass->params, &pass->image, &pass->target))
        return false;
    if (!find_fbo_format(pass))
        return false;
    if (!pass_fix_frames(pass))
        return false;
    if (rr->have_frame_index)
        pl_dispatch_set_frame_index(rr->dp, rr->frame_index);
    else
        pl_dispatch_reset_frame(rr->dp);
    memset(rr->fbo_used, 0, sizeof(rr->fbo_used));
    return true;
}

#define MEASURED_SLOTS 4

// keep the FBOs of all pending measured images out of the pool
static void reserve_measured_fbos(pl_renderer rr)
{
    for (int m = 0; m < MEASURED_SLOTS; m++) {
        if (!rr->measured[m].valid)
            continue;
        if (rr->measure_seq - rr->measured[m].seq >= MEASURED_SLOTS) { // never consumed
            rr->measured[m].valid = false;
            continue;
        }
        for (int i = 0; i < MAX_FBOS; i++) {
            if (rr->fbos[i] == rr->measured[m].tex)
                rr->fbo_used[i] = true;
        }
    }
}

// Does the image of this render call equal one pl_renderer_measure_peak has processed? Compares
// what determines the stored image (planes, crop, representation, colour space, grain, frame
// index, render parameters by value), not just pointers.
static struct measured_image *measured_find(pl_renderer rr, const struct pl_frame *image,
                                            const struct pl_render_params *params)
{
    if (!rr->have_peak_override)
        return NULL;
    for (int m = 0; m < MEASURED_SLOTS; m++) {
        struct measured_image *mi = &rr->measured[m];
        if (!mi->valid || mi->params_ptr != params || memcmp(&mi->params, params, sizeof(*params)))
            continue;
        if (rr->have_frame_index && mi->frame_index != rr->frame_index)
            continue;
        bool same = true;
        for (int i = 0; i < 4; i++) {
            pl_tex t = i < image->num_planes ? image->planes[i].texture : NULL;
            same &= t == mi->planes[i];
        }
        same = same && !memcmp(&mi->crop, &image->crop, sizeof(mi->crop)) &&
               pl_color_repr_equal(&mi->in_repr, &image->repr) &&
               pl_color_space_equal(&mi->in_color, &image->color) &&
               mi->grain.type == image->film_grain.type &&
               (!mi->grain.type || !memcmp(&mi->grain, &image->film_grain, sizeof(mi->grain)));
        if (same)
            return mi;
    }
    return NULL;
}

bool pl_render_image(pl_renderer rr, const struct pl_frame *pimage,
                     const struct pl_frame *ptarget, const struct pl_render_params *params)
{
    params = params ? params : &pl_render_default_params;
    if (!pimage || !ptarget) {
        PLB_ERR(rr->log, "pl_render_image: image and target are required");
        return false;
    }
    struct pass_state pass = {
        .rr = rr,
        .params = params,
        .image = *pimage,
        .target = *ptarget,
    };
    if (!pass_begin(&pass))
        return false;
    if (!pl_rect_w(pass.dst_rect) || !pl_rect_h(pass.dst_rect))
        return true; // empty crop
    rr->info_params = params;
    rr->info_index = 0;
    pl_dispatch_callback(rr->dp, rr, params->info_callback ? renderer_info_cb : NULL);
    bool ok;
    reserve_measured_fbos(rr);
    struct measured_image *mi = measured_find(rr, pimage, params);
    if (mi) {
        pass.from_measured = true;
        pass.tex = mi->tex;
        pass.w = mi->w;
        pass.h = mi->h;
        pass.comps = mi->comps;
        pass.rect = mi->rect;
        pass.ref_rect = mi->ref_rect;
        pass.repr = mi->repr;
        pass.color = mi->color;
        mi->valid = false; // single use (its FBO stays reserved for this frame)
        ok = true;
    } else {
        ok = pass_read_image(&pass);
    }
    ok = ok && pass_scale_main(&pass) && pass_convert_colors(&pass) && pass_output_target(&pass);
    rr->info_params = NULL;
    free_views(rr);
    if (!ok) {
        PLB_ERR(rr->log, "Failed rendering image!");
        pl_dispatch_abort(rr->dp, &pass.sh);
    }
    return ok;
}

// ---------------------------------------------------------------------------------------
// pl_render_image_mix (renderer.c:3477-3508, :3612-4025)
// ---------------------------------------------------------------------------------------

// the frame on screen at the vsync: the last one whose timestamp is not in the future
const struct pl_frame *pl_frame_mix_current(const struct pl_frame_mix *mix)
{
    int idx = -1;
    while (idx + 1 < mix->num_frames && !(mix->timestamps[idx + 1] > 0.0f))
        idx++;
    return idx < 0 ? NULL : mix->frames[idx];
}

// the frame closest to the vsync (timestamps are sorted: stop at the first one farther away)
const struct pl_frame *pl_frame_mix_nearest(const struct pl_frame_mix *mix)
{
    if (mix->num_frames <= 0)
        return NULL;
    int best = 0;
    for (int i = 1; i < mix->num_frames; i++) {
        if (!(fabsf(mix->timestamps[i]) < fabsf(mix->timestamps[best])))
            break;
        best = i;
    }
    return mix->frames[best];
}

// identity of everything in the parameters that shapes a cached frame (the reference hashes the
// pointed-to structures, :3510-3560; here: the struct itself plus the structures it points to)
static uint64_t params_hash(const struct pl_render_params *p)
{
    uint64_t h = 0xcbf29ce484222325ull;
#define MIX(ptr, n) do { const uint8_t *b_ = (const uint8_t *) (ptr); \
        for (size_t i_ = 0; i_ < (n); i_++) h = (h ^ b_[i_]) * 0x100000001b3ull; } while (0)
    MIX(p, sizeof(*p));
    if (p->upscaler) MIX(p->upscaler, sizeof(*p->upscaler));
    if (p->downscaler) MIX(p->downscaler, sizeof(*p->downscaler));
    if (p->plane_upscaler) MIX(p->plane_upscaler, sizeof(*p->plane_upscaler));
    if (p->deband_params) MIX(p->deband_params, sizeof(*p->deband_params));
    if (p->sigmoid_params) MIX(p->sigmoid_params, sizeof(*p->sigmoid_params));
    if (p->color_map_params) MIX(p->color_map_params, sizeof(*p->color_map_params));
    if (p->peak_detect_params) MIX(p->peak_detect_params, sizeof(*p->peak_detect_params));
#undef MIX
    return h;
}

extern float plb_filter_radius_bound(const struct pl_filter_config *c);
extern bool plb_shader_mix(pl_shader sh, int n, const pl_tex *texs, const float *weights, int w, int h,
                           const struct pl_color_space *lin_csp);

// Contribution of frame `i` to the vsync (renderer.c:3689-3738). mixer == NULL: only the
// reference image counts (zero-order hold); oversampling: the visible fraction of the vsync
// interval; any other filter: its value at the frame's timestamp. Returns false for frames
// that do not take part at all.
static bool mix_frame_weight(const struct pl_frame_mix *mix, int i, const struct pl_filter_config *mixer,
                             bool is_ref, float *out)
{
    const float vsync = mix->vsync_duration;
    const float start = mix->timestamps[i];
    if (!mixer) {
        *out = 1.0;
        return is_ref;
    }
    if (!mixer->kernel || mixer->kernel == &pl_filter_function_oversample) {
        const float stop = i + 1 < mix->num_frames ? mix->timestamps[i + 1] : INFINITY;
        if (start > vsync || stop < 0.0)
            return false; // no intersection with the vsync interval
        const float a = PLB_MAX(start, 0.0), b = PLB_MIN(stop, vsync);
        float w = (b - a) / vsync;
        if (mixer->kernel && w < mixer->kernel->params[0])
            w = 0.0; // below the oversampling threshold
        *out = w;
        return true;
    }
    if (fabsf(start) >= plb_filter_radius_bound(mixer))
        return false;
    *out = pl_filter_sample(mixer, start);
    return true;
}

bool pl_render_image_mix(pl_renderer rr, const struct pl_frame_mix *images,
                         const struct pl_frame *ptarget, const struct pl_render_params *params)
{
    params = params ? params : &pl_render_default_params;
    if (!images || !ptarget) {
        PLB_ERR(rr->log, "pl_render_image_mix: frame mix and target are required");
        return false;
    }
    if (!images->num_frames) {
        PLB_ERR(rr->log, "pl_render_image_mix: rendering without a source image (target clearing "
                "only) is outside the CUDA hot path");
        return false;
    }
    if (!(images->vsync_duration > 0.0)) {
        PLB_ERR(rr->log, "pl_render_image_mix: vsync_duration must be positive");
        return false;
    }
    for (int i = 0; i < images->num_frames - 1; i++) {
        if (!(images->timestamps[i] <= images->timestamps[i + 1])) {
            PLB_ERR(rr->log, "pl_render_image_mix: timestamps must be sorted");
            return false;
        }
    }
    const struct pl_frame *refimg = pl_frame_mix_nearest(images);
    if (rr->errors & PL_RENDER_ERR_FRAME_MIXING)
        return pl_render_image(rr, refimg, ptarget, params);

    struct pass_state pass = { .rr = rr, .params = params, .image = *refimg, .target = *ptarget };
    if (!pass_begin(&pass))
        return false;
    const struct pl_frame *target = &pass.target;
    const int out_w = abs(pl_rect_w(pass.dst_rect)), out_h = abs(pl_rect_h(pass.dst_rect));
    if (!out_w || !out_h)
        return true;
    const uint64_t phash = params_hash(params);

    for (int i = 0; i < rr->num_frames; i++)
        rr->frames[i].evict = true;

    // blur the frame mixer according to the vsync ratio (source / display), :3661-3674
    struct pl_filter_config mixer = {0};
    if (params->frame_mixer) {
        mixer = *params->frame_mixer;
        mixer.blur = PLB_DEF(mixer.blur, 1.0);
        for (int i = 1; i < images->num_frames; i++) {
            if (images->timestamps[i] >= 0.0 && images->timestamps[i - 1] < 0) {
                const float frame_dur = images->timestamps[i] - images->timestamps[i - 1];
                const float sample_dur = PLB_MAX(frame_dur, images->vsync_duration);
                if (sample_dur > 1.0f && !params->skip_anti_aliasing)
                    mixer.blur *= sample_dur;
                break;
            }
        }
    }

    pl_tex texs[PLB_MAX_MIX];
    float weights[PLB_MAX_MIX];
    int fidx = 0;
    float wsum = 0.0;
    bool single_frame = !params->frame_mixer || images->num_frames == 1;
retry:
    for (int i = 0; i < images->num_frames; i++) {
        const uint64_t sig = images->signatures[i];
        const struct pl_frame *img = images->frames[i];
        float weight;
        if (!mix_frame_weight(images, i, single_frame ? NULL : &mixer, img == refimg, &weight))
            continue;

        struct cached_frame *f = NULL;
        for (int j = 0; j < rr->num_frames; j++) {
            if (rr->frames[j].signature == sig) {
                f = &rr->frames[j];
                f->evict = false;
                break;
            }
        }
        // frames with negligible contributions are skipped (never the reference image)
        if (fabsf(weight) <= 1e-3 && img != refimg)
            continue;
        if (fidx == PLB_MAX_MIX) {
            PLB_ERR(rr->log, "pl_render_image_mix: more than %d contributing frames", PLB_MAX_MIX);
            goto fail;
        }
        if (!f) {
            if (rr->num_frames == (int) PLB_ARRAY_SIZE(rr->frames)) {
                PLB_ERR(rr->log, "pl_render_image_mix: frame cache full");
                goto fail;
            }
            f = &rr->frames[rr->num_frames++];
            memset(f, 0, sizeof(*f));
            f->signature = sig;
        }
        const bool can_reuse = f->tex && f->tex->params.w == out_w && f->tex->params.h == out_h &&
                               pl_rect2d_eq(f->crop, img->crop) && f->params_hash == phash &&
                               pl_color_space_equal(&f->color, &target->color);
        if (!can_reuse) {
            if (!pl_tex_recreate(rr->gpu, &f->tex, pl_tex_params(
                    .w = out_w, .h = out_h, .format = rr->fbofmt[4],
                    .sampleable = true, .renderable = true, .storable = true))) {
                PLB_ERR(rr->log, "Could not create intermediate texture for frame mixing.. disabling!");
                goto fail;
            }
            // render this frame up to (not including) the output stage, :3822-3853
            struct pass_state ip = { .rr = rr, .params = params, .image = *img, .target = *ptarget };
            bool ok = pass_begin(&ip) && pass_read_image(&ip) && pass_scale_main(&ip) &&
                      pass_convert_colors(&ip);
            if (ok) {
                pl_shader ish = img_sh(&ip);
                ip.sh = NULL;
                pl_shader_set_alpha(ish, &ip.repr, PL_ALPHA_PREMULTIPLIED); // for frame mixing
                ok = ip.w == out_w && ip.h == out_h &&
                     pl_dispatch_finish(rr->dp, pl_dispatch_params( .shader = &ish, .target = f->tex ));
                if (ish)
                    pl_dispatch_abort(rr->dp, &ish);
            } else {
                pl_dispatch_abort(rr->dp, &ip.sh);
            }
            free_views(rr);
            if (!ok)
                goto fail;
            f->params_hash = phash;
            f->crop = img->crop;
            f->color = ip.color;
            f->comps = ip.comps;
        }
        texs[fidx] = f->tex;
        weights[fidx] = weight;
        wsum += weight;
        fidx++;
    }

    // evict the frames we do not need any more
    for (int i = 0; i < rr->num_frames;) {
        if (rr->frames[i].evict) {
            pl_tex_destroy(rr->gpu, &rr->frames[i].tex);
            rr->frames[i] = rr->frames[--rr->num_frames];
        } else {
            i++;
        }
    }

    if (!fidx) { // no frame intersects the vsync: zero-order hold
        single_frame = true;
        goto retry;
    }

    // sample and mix the output colour, :3906-3990
    memset(rr->fbo_used, 0, sizeof(rr->fbo_used));
    const bool mixing = fidx > 1;
    struct pl_color_space mix_csp = target->color;
    if (mixing)
        mix_csp.transfer = PL_COLOR_TRC_LINEAR;
    for (int i = 0; i < fidx; i++)
        weights[i] = mixing ? weights[i] / wsum : 1.0f;
    pl_shader sh = begin(&pass);
    if (!plb_shader_mix(sh, fidx, texs, weights, out_w, out_h, mixing ? &target->color : NULL)) {
        pl_dispatch_abort(rr->dp, &sh);
        goto fail;
    }
    pass.sh = sh;
    pass.tex = NULL;
    pass.w = out_w;
    pass.h = out_h;
    pass.comps = 3;
    pass.color = target->color;
    pass.rect = (pl_rect2df) { 0, 0, out_w, out_h };
    pass.repr = (struct pl_color_repr) { .sys = PL_COLOR_SYSTEM_RGB, .levels = PL_COLOR_LEVELS_FULL,
                                         .alpha = PL_ALPHA_NONE };
    // re-encode to the target transfer (in practice: delinearise)
    if (!pl_color_space_equal(&mix_csp, &pass.color))
        pl_shader_color_map_ex(sh, NULL, pl_color_map_args( .src = mix_csp, .dst = pass.color ));
    if (!pass_output_target(&pass)) {
        pl_dispatch_abort(rr->dp, &pass.sh);
        return false;
    }
    return true;

fail:
    PLB_ERR(rr->log, "Could not render image for frame mixing.. disabling!");
    rr->errors |= PL_RENDER_ERR_FRAME_MIXING;
    return pl_render_image(rr, refimg, ptarget, params);
}

// Phase 1 of the frame-parallel scheme: run the path up to (and including) the peak detection
// tap, keep only the raw statistics
static bool measure_peak(pl_renderer rr, const struct pl_frame *pimage,
                         const struct pl_render_params *params, void *out_peak_buf, bool to_device)
{
    params = params ? params : &pl_render_default_params;
    if (!params->peak_detect_params) {
        PLB_ERR(rr->log, "pl_renderer_measure_peak: peak_detect_params required");
        return false;
    }
    if (!pimage || !pimage->num_planes || !pimage->planes[0].texture || !out_peak_buf) {
        PLB_ERR(rr->log, "pl_renderer_measure_peak: image with at least one plane and an output "
                "buffer are required");
        return false;
    }
    for (int i = 0; i < pimage->num_planes; i++) {
        if (!pimage->planes[i].texture) {
            PLB_ERR(rr->log, "pl_renderer_measure_peak: plane %d has no texture", i);
            return false;
        }
    }
    // a dummy target of the same size/space as the image: only its colour space matters
    struct pl_frame target = {
        .num_planes = 1,
        .planes = {{ .texture = pimage->planes[0].texture, .components = 3,
                     .component_mapping = {0, 1, 2} }},
        .repr = { .sys = PL_COLOR_SYSTEM_RGB, .levels = PL_COLOR_LEVELS_FULL },
        .color = pl_color_space_bt709,
    };
    struct pass_state pass = { .rr = rr, .params = params, .image = *pimage, .target = target };
    const bool had_override = rr->have_peak_override;
    rr->have_peak_override = false;
    bool ok = false;
    struct pl_render_params hacked = *params;
    hacked.plane_upscaler = PLB_DEF(params->plane_upscaler, params->upscaler);
    hacked.plane_downscaler = PLB_DEF(params->plane_downscaler, params->downscaler);
    hacked.upscaler = hacked.downscaler = NULL;
    pass.params = &hacked;
    if (!find_fbo_format(&pass))
        goto done;
    // skip the rect fixing against the dummy target: measure on the source grid
    pass.src_ref = frame_ref(&pass.image);
    {
        pl_tex ref = pass.image.planes[pass.src_ref].texture;
        pl_rect2df *crop = &pass.image.crop;
        if ((!crop->x0 && !crop->x1) || (!crop->y0 && !crop->y1))
            *crop = (pl_rect2df) { 0, 0, ref->params.w, ref->params.h };
        pass.dst_rect = (pl_rect2d) { 0, 0, roundf(pl_rect_w(*crop)), roundf(pl_rect_h(*crop)) };
    }
    fix_frame(&pass.image);
    pl_color_space_infer_map(&pass.image.color, &pass.target.color);
    if (pass.image.repr.alpha == PL_ALPHA_UNKNOWN)
        pass.image.repr.alpha = PL_ALPHA_INDEPENDENT;
    if (rr->have_frame_index)
        pl_dispatch_set_frame_index(rr->dp, rr->frame_index);
    else
        pl_dispatch_reset_frame(rr->dp);
    memset(rr->fbo_used, 0, sizeof(rr->fbo_used));
    rr->measure_seq++;
    reserve_measured_fbos(rr); // images measured earlier and not rendered yet stay intact
    rr->info_params = params; // the measuring passes are reported like pl_render_image's
    rr->info_index = 0;
    pl_dispatch_callback(rr->dp, rr, params->info_callback ? renderer_info_cb : NULL);
    if (!pass_read_image(&pass))
        goto done;
    pl_reset_detected_peak(rr->tone_map_state);
    if (!pl_shader_detect_peak(img_sh(&pass), pass.color, &rr->tone_map_state,
                               params->peak_detect_params))
        goto done;
    // statistics + the image itself (rgba16hf FBO, exactly what the sequential same-frame flow
    // stores between its two passes): the following pl_render_image continues from it
    ok = img_tex(&pass) != NULL;
    if (ok) {
        struct plb_peak_state *pk = &rr->tone_map_state->u.cmap.peak;
        ok = to_device ? plb_dev_copy(rr->gpu, out_peak_buf, pk->dev_buf, PL_PEAK_BUF_SIZE)
                       : plb_dev_download(rr->gpu, out_peak_buf, pk->dev_buf, PL_PEAK_BUF_SIZE);
        pk->pending = false;
    }
    if (ok) {
        // the slot of the oldest (or an invalid) entry
        struct measured_image *mi = &rr->measured[0];
        for (int m = 0; m < MEASURED_SLOTS; m++) {
            if (!rr->measured[m].valid) {
                mi = &rr->measured[m];
                break;
            }
            if (rr->measured[m].seq < mi->seq)
                mi = &rr->measured[m];
        }
        mi->valid = true;
        mi->seq = rr->measure_seq;
        mi->tex = pass.tex;
        for (int i = 0; i < 4; i++)
            mi->planes[i] = i < pimage->num_planes ? pimage->planes[i].texture : NULL;
        mi->params = *params;
        mi->params_ptr = params;
        mi->crop = pimage->crop;
        mi->in_repr = pimage->repr;
        mi->in_color = pimage->color;
        mi->grain = pimage->film_grain;
        mi->frame_index = rr->frame_index;
        mi->w = pass.w;
        mi->h = pass.h;
        mi->comps = pass.comps;
        mi->rect = pass.rect;
        mi->ref_rect = pass.ref_rect;
        mi->repr = pass.repr;
        mi->color = pass.color;
    }
done:
    rr->info_params = NULL;
    pl_dispatch_abort(rr->dp, &pass.sh);
    free_views(rr);
    rr->have_peak_override = had_override;
    return ok;
}

bool pl_renderer_measure_peak(pl_renderer rr, const struct pl_frame *pimage,
                              const struct pl_render_params *params, void *out_peak_buf)
{
    return measure_peak(rr, pimage, params, out_peak_buf, false);
}

// The same, but the statistics stay on the device: `dev_peak_buf` (PL_PEAK_BUF_SIZE bytes of
// device memory) is filled by an asynchronous copy on the gpu's stream, the host does not wait.
bool pl_renderer_measure_peak_dev(pl_renderer rr, const struct pl_frame *pimage,
                                  const struct pl_render_params *params, void *dev_peak_buf)
{
    return measure_peak(rr, pimage, params, dev_peak_buf, true);
}


// ---------------------------------------------------------------------------------------
// pl_frame_clear_rgba (renderer.h:675, src/renderer.c:4169-4192): an sRGB colour is taken to
// the frame's colour space (translate_srgb_color :2555-2584: same transfer curve for the
// BT.1886 / sRGB / gamma 2.2 family, else infinite-contrast sRGB; relative colorimetric
// primaries change), encoded with the inverse of pl_color_repr_decode and written to every
// plane with pl_tex_clear. Host arithmetic on one colour; the transfer functions below are this
// library's own statement of pl_shader_linearize / delinearize for a single float triple.
// ---------------------------------------------------------------------------------------

struct trc_ctx { float lo, hi; }; // nominal black / peak of the space, relative to SDR white

static float pq_to_linear(float e)
{
    const double m1 = 2610.0 / 16384, m2 = 2523.0 / 32, c1 = 3424.0 / 4096, c2 = 2413.0 / 128, c3 = 2392.0 / 128;
    const double p = pow(fmax(e, 0.0), 1.0 / m2);
    return (float) (pow(fmax(p - c1, 0.0) / (c2 - c3 * p), 1.0 / m1) * 10000.0 / PL_COLOR_SDR_WHITE);
}

static float linear_to_pq(float l)
{
    const double m1 = 2610.0 / 16384, m2 = 2523.0 / 32, c1 = 3424.0 / 4096, c2 = 2413.0 / 128, c3 = 2392.0 / 128;
    const double y = pow(fmax(l, 0.0) * PL_COLOR_SDR_WHITE / 10000.0, m1);
    return (float) pow((c1 + c2 * y) / (1.0 + c3 * y), m2);
}

static void clear_trc(const struct pl_color_space *csp, float c[3], bool to_linear)
{
    struct trc_ctx t;
    pl_color_space_nominal_luma_ex(pl_nominal_luma_params(
        .color = csp, .metadata = PL_HDR_METADATA_HDR10, .scaling = PL_HDR_NORM,
        .out_min = &t.lo, .out_max = &t.hi));
    const float span = t.hi - t.lo;
    for (int k = 0; k < 3; k++) {
        float x = c[k];
        switch (csp->transfer) {
        case PL_COLOR_TRC_LINEAR:
            break;
        case PL_COLOR_TRC_PQ:
            x = to_linear ? pq_to_linear(x) : linear_to_pq(x);
            break;
        case PL_COLOR_TRC_BT_1886: {
            const float lb = powf(t.lo, 1 / 2.4f), lw = powf(t.hi, 1 / 2.4f);
            const float a = powf(lw - lb, 2.4f), b = lb / (lw - lb);
            x = to_linear ? a * powf(fmaxf(x, 0) + b, 2.4f) : powf(fmaxf(x, 0) / a, 1 / 2.4f) - b;
            break;
        }
        case PL_COLOR_TRC_SRGB:
            if (to_linear) {
                x = fmaxf(x, 0);
                x = x > 0.04045f ? powf((x + 0.055f) / 1.055f, 2.4f) : x / 12.92f;
                x = span * x + t.lo;
            } else {
                x = fmaxf((x - t.lo) / span, 0);
                x = x >= 0.0031308f ? 1.055f * powf(x, 1 / 2.4f) - 0.055f : 12.92f * x;
            }
            break;
        default: { // pure power curves and the remaining spaces: gamma 2.2 unless a plain gamma
            float g = 2.2f;
            switch (csp->transfer) {
            case PL_COLOR_TRC_GAMMA18: g = 1.8f; break;
            case PL_COLOR_TRC_GAMMA20: g = 2.0f; break;
            case PL_COLOR_TRC_GAMMA24: g = 2.4f; break;
            case PL_COLOR_TRC_GAMMA26: g = 2.6f; break;
            case PL_COLOR_TRC_GAMMA28: g = 2.8f; break;
            default: break;
            }
            x = to_linear ? span * powf(fmaxf(x, 0), g) + t.lo : powf(fmaxf((x - t.lo) / span, 0), 1 / g);
            break;
        }
        }
        c[k] = x;
    }
}

void pl_frame_clear_rgba(pl_gpu gpu, const struct pl_frame *frame, const float rgba[4])
{
    if (!frame || !rgba)
        return;
    struct pl_color_space dst = frame->color;
    pl_color_space_infer(&dst);
    struct pl_color_space srgb = pl_color_space_srgb;
    if (dst.transfer == PL_COLOR_TRC_BT_1886 || dst.transfer == PL_COLOR_TRC_SRGB ||
        dst.transfer == PL_COLOR_TRC_GAMMA22) {
        srgb.transfer = dst.transfer; // re-use the curve: no small adaptations
        srgb.hdr.min_luma = dst.hdr.min_luma;
    } else {
        srgb.hdr.min_luma = PL_COLOR_HDR_BLACK;
    }
    pl_color_space_infer(&srgb);
    float enc[3] = { rgba[0], rgba[1], rgba[2] };
    clear_trc(&srgb, enc, true);
    const pl_matrix3x3 prim = pl_get_color_mapping_matrix(pl_raw_primaries_get(srgb.primaries),
                                                          pl_raw_primaries_get(dst.primaries),
                                                          PL_INTENT_RELATIVE_COLORIMETRIC);
    pl_matrix3x3_apply(&prim, enc);
    clear_trc(&dst, enc, false);

    struct pl_color_repr repr = frame->repr;
    pl_transform3x3 tr = pl_color_repr_decode(&repr, NULL);
    pl_transform3x3_invert(&tr);
    pl_transform3x3_apply(&tr, enc);

    const float mult = frame->repr.alpha == PL_ALPHA_PREMULTIPLIED ? rgba[3] : 1.0f;
    for (int p = 0; p < frame->num_planes; p++) {
        const struct pl_plane *plane = &frame->planes[p];
        float clear[4] = { 0.0f, 0.0f, 0.0f, rgba[3] };
        for (int c = 0; c < plane->components; c++) {
            const int ch = plane->component_mapping[c];
            if (ch >= 0 && ch < 3)
                clear[c] = mult * enc[ch];
        }
        if (plane->texture)
            pl_tex_clear(gpu, plane->texture, clear);
    }
}
