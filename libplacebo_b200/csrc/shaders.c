// pl_shader on the CUDA backend: every pl_shader_* entry point of the hot path RECORDS a typed
// op (with all parameters derived on the host exactly like the reference derives its GLSL
// constants) instead of emitting GLSL text. Same call protocol as the reference: one
// sampling/source call first, colour ops after it, failures latch in pl_shader_is_failed.
//
//   pl_shader_sample_polar   src/shaders/sampling.c:587-912  (host part :596-631, :661, :720)
//   pl_shader_sample_ortho2  sampling.c:950-1104
//   pl_shader_sample_direct/nearest/bilinear  sampling.c:277-340
//   pl_shader_deband         sampling.c:183-275
//   pl_shader_decode_color   src/shaders/colorspace.c:275-460, encode :462-573
//   pl_shader_linearize/delinearize :589-847, sigmoid :851-894
//   pl_shader_dither         src/shaders/dithering.c:109-274
#define _GNU_SOURCE
#include "shader_priv.h"

#include <math.h>

// ---------------------------------------------------------------------------------------
// shader object plumbing (src/shaders.c:60-130, :909-963)
// ---------------------------------------------------------------------------------------

pl_shader pl_shader_alloc(pl_log log, const struct pl_shader_params *params)
{
    struct pl_shader_t *sh = calloc(1, sizeof(*sh));
    sh->log = log;
    pl_shader_reset(sh, params);
    return sh;
}

static void shader_unref_objs(pl_shader sh)
{
    for (int i = 0; i < sh->num_objs; i++) {
        pl_shader_obj obj = sh->objs[i];
        pl_shader_obj_destroy(&obj);
    }
    sh->num_objs = 0;
}

void pl_shader_free(pl_shader *psh)
{
    if (!psh || !*psh)
        return;
    shader_unref_objs(*psh);
    free(*psh);
    *psh = NULL;
}

void pl_shader_reset(pl_shader sh, const struct pl_shader_params *params)
{
    pl_log log = sh->log;
    pl_dispatch dp = sh->dp;
    shader_unref_objs(sh);
    memset(sh, 0, sizeof(*sh));
    sh->log = log;
    sh->dp = dp;
    if (params)
        sh->params = *params;
    sh->px.init[3] = 1.0f;
    for (int k = 0; k < 4; k++) {
        sh->polar.src_map[k] = k;
        sh->ortho.src_map[k] = k;
        sh->mix.src_map[k] = k;
        sh->deint.src_map[k] = k;
    }
}

bool pl_shader_is_failed(const pl_shader sh) { return sh->failed; }
bool pl_shader_is_compute(const pl_shader sh) { return sh->is_compute; }

bool pl_shader_output_size(const pl_shader sh, int *w, int *h)
{
    if (!sh->out_w || !sh->out_h)
        return false;
    *w = sh->out_w;
    *h = sh->out_h;
    return true;
}

void plb_shader_fail(pl_shader sh, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    PLB_ERR(sh->log, "%s", buf);
    sh->failed = true;
}

void pl_shader_obj_destroy(pl_shader_obj *ptr)
{
    pl_shader_obj obj = ptr ? *ptr : NULL;
    if (!obj)
        return;
    *ptr = NULL;
    if (--obj->refcount > 0)
        return;
    pl_gpu gpu = obj->gpu;
    switch (obj->type) {
    case PLB_OBJ_SAMPLER:
        for (struct plb_sampler_obj *s = &obj->u.sampler; s;) {
            pl_filter_free(&s->filter);
            plb_dev_free(gpu, s->dev_lut);
            free(s->host_lut);
            struct plb_sampler_obj *next = s->pass2;
            if (s != &obj->u.sampler)
                free(s);
            s = next;
        }
        break;
    case PLB_OBJ_COLOR_MAP:
        plb_dev_free(gpu, obj->u.cmap.dev_tone_lut);
        plb_dev_free(gpu, obj->u.cmap.dev_gamut_lut);
        plb_dev_free(gpu, obj->u.cmap.peak.dev_buf);
        break;
    case PLB_OBJ_DITHER:
        plb_dev_free(gpu, obj->u.dither.dev_lut);
        break;
    case PLB_OBJ_LUT:
        plb_dev_free(gpu, obj->u.lut.dev_lut);
        break;
    case PLB_OBJ_GRAIN: {
        extern void plb_grain_obj_free(pl_gpu gpu, struct plb_grain_obj *g);
        plb_grain_obj_free(gpu, obj->u.grain);
        break;
    }
    default:
        break;
    }
    free(obj);
}

// SH_OBJ: lazily create a state object of the given type and keep it alive with the shader
void *plb_shader_obj(pl_shader sh, pl_shader_obj *ptr, enum plb_obj_type type)
{
    if (!ptr) {
        plb_shader_fail(sh, "shader state object required but none given");
        return NULL;
    }
    pl_shader_obj obj = *ptr;
    if (obj && obj->gpu != sh->params.gpu) {
        plb_shader_fail(sh, "Passed pl_shader_obj belongs to different GPU!");
        return NULL;
    }
    if (obj && obj->type != type) {
        plb_shader_fail(sh, "Passed pl_shader_obj of wrong type! Shader objects must always "
                        "be used with the same type of shader.");
        return NULL;
    }
    if (!obj) {
        obj = calloc(1, sizeof(*obj));
        obj->type = type;
        obj->gpu = sh->params.gpu;
        obj->refcount = 1;
        *ptr = obj;
    }
    if (sh->num_objs < (int) PLB_ARRAY_SIZE(sh->objs)) {
        obj->refcount++;
        sh->objs[sh->num_objs++] = obj;
    }
    return &obj->u;
}

struct plb_color_op *plb_shader_push_op(pl_shader sh, int kind)
{
    if (sh->px.prog.n >= PLB_MAX_OPS) {
        plb_shader_fail(sh, "too many colour operations in one shader (max %d)", PLB_MAX_OPS);
        return NULL;
    }
    struct plb_color_op *op = &sh->px.prog.ops[sh->px.prog.n++];
    memset(op, 0, sizeof(*op));
    op->kind = kind;
    // friendly step names, worded like the reference's sh_describe() calls
    static const char *const names[] = {
        [PLB_OP_MATRIX] = "color matrix",
        [PLB_OP_LINEARIZE] = "linearization",
        [PLB_OP_DELINEARIZE] = "delinearization",
        [PLB_OP_DECODE_SPECIAL] = "color decoding",
        [PLB_OP_ENCODE_SPECIAL] = "color encoding",
        [PLB_OP_SIGMOID] = "sigmoidization",
        [PLB_OP_UNSIGMOID] = "unsigmoidization",
        [PLB_OP_COLOR_MAP] = "tone / gamut mapping",
        [PLB_OP_DITHER] = "dithering",
        [PLB_OP_DOVI_RESHAPE] = "reshaping",
        [PLB_OP_DOVI_LMS] = "color decoding",
        [PLB_OP_PEAK] = "peak detection",
        [PLB_OP_CUSTOM_LUT] = "custom LUT",
    };
    if (kind >= 0 && kind < (int) PLB_ARRAY_SIZE(names) && names[kind])
        plb_shader_describe(sh, names[kind]);
    return op;
}

void plb_shader_use_tex(pl_shader sh, pl_tex tex)
{
    for (int i = 0; i < sh->num_texs; i++) {
        if (sh->texs[i] == tex)
            return;
    }
    if (sh->num_texs < (int) PLB_ARRAY_SIZE(sh->texs))
        sh->texs[sh->num_texs++] = tex;
}

void plb_shader_describe(pl_shader sh, const char *desc)
{
    if (sh->num_steps < (int) PLB_ARRAY_SIZE(sh->steps))
        sh->steps[sh->num_steps++] = desc;
}

// sh_require(sh, PL_SHADER_SIG_COLOR, 0, 0)
bool plb_shader_require_color(pl_shader sh, const char *who)
{
    if (sh->failed)
        return false;
    if (sh->src == PLB_SRC_NONE) {
        plb_shader_fail(sh, "%s: shader has no input colour yet (call a sampling shader "
                        "first)", who);
        return false;
    }
    return true;
}

static bool require_input_none(pl_shader sh, int w, int h, const char *who)
{
    if (sh->failed)
        return false;
    if (sh->src != PLB_SRC_NONE) {
        plb_shader_fail(sh, "%s: illegal sequence of shader operations (shader already has "
                        "a source stage)", who);
        return false;
    }
    if ((sh->out_w && w && sh->out_w != w) || (sh->out_h && h && sh->out_h != h)) {
        plb_shader_fail(sh, "%s: incompatible output size requirements %dx%d and %dx%d", who,
                        sh->out_w, sh->out_h, w, h);
        return false;
    }
    sh->out_w = PLB_DEF(sh->out_w, w);
    sh->out_h = PLB_DEF(sh->out_h, h);
    return true;
}

// sh_try_compute (src/shaders.c:214-285), minus the shared-memory accounting
bool plb_shader_try_compute(pl_shader sh, int bw, int bh, bool flex)
{
    if (bw * bh > 1024) {
        if (!flex)
            return false;
        bw = PLB_MIN(bw, 1024);
        bh = 1024 / bw;
    }
    // not a compute shader yet, or a flexible geometry meets a rigid one: take the new one
    if (!sh->is_compute || (sh->wg_flexible && !flex)) {
        sh->wg_w = bw;
        sh->wg_h = bh;
        sh->is_compute = true;
        sh->wg_flexible = flex;
        return true;
    }
    if (sh->wg_flexible && flex) {
        sh->wg_w = PLB_MAX(sh->wg_w, bw);
        sh->wg_h = PLB_MAX(sh->wg_h, bh);
        return true;
    }
    // existing rigid geometry: a second rigid request must match, a flexible one adapts
    if (!flex && (bw != sh->wg_w || bh != sh->wg_h))
        return false;
    return true;
}

// ---------------------------------------------------------------------------------------
// setup_src (sampling.c:45-181)
// ---------------------------------------------------------------------------------------

struct src_info {
    float src_w, src_h;
    int out_w, out_h;
    float rx, ry, scale;
    uint8_t mask;
    pl_rect2df rect;
    struct plb_tex_desc tex;
};

static bool setup_src(pl_shader sh, const struct pl_sample_src *src, struct src_info *o,
                      bool resizable, bool need_linear, const char *who)
{
    if (!src || !src->tex) {
        plb_shader_fail(sh, "%s: pl_sample_src.tex is required on the CUDA backend (the "
                        "sampler2D signature cannot be honoured without GLSL)", who);
        return false;
    }
    pl_tex tex = src->tex;
    pl_fmt fmt = tex->params.format;
    if (pl_tex_params_dimension(tex->params) != 2) {
        plb_shader_fail(sh, "%s: not a 2D texture!", who);
        return false;
    }
    if (need_linear && !(fmt->caps & PL_FMT_CAP_LINEAR)) {
        plb_shader_fail(sh, "Trying to use a shader that requires linear sampling with a "
                        "texture whose format (%s) does not support PL_FMT_CAP_LINEAR",
                        fmt->name);
        return false;
    }
    o->src_w = pl_rect_w(src->rect);
    o->src_h = pl_rect_h(src->rect);
    o->src_w = PLB_DEF(o->src_w, tex->params.w);
    o->src_h = PLB_DEF(o->src_h, tex->params.h);
    o->out_w = PLB_DEF(src->new_w, (int) roundf(fabs(o->src_w)));
    o->out_h = PLB_DEF(src->new_h, (int) roundf(fabs(o->src_h)));
    o->rx = o->out_w / fabs(o->src_w);
    o->ry = o->out_h / fabs(o->src_h);
    o->scale = PLB_DEF(src->scale, 1.0);
    if (fmt->type == PL_FMT_UINT)
        o->scale *= 1.0 / ((1ull << fmt->component_depth[0]) - 1);
    uint8_t tex_mask = (1 << fmt->num_components) - 1;
    uint8_t src_mask = src->component_mask;
    if (!src_mask)
        src_mask = (1 << PLB_DEF(src->components, 4)) - 1;
    o->mask = tex_mask & src_mask;
    o->rect = (pl_rect2df) {
        src->rect.x0, src->rect.y0, src->rect.x0 + o->src_w, src->rect.y0 + o->src_h,
    };
    o->tex = plb_tex_priv(tex)->d;
    if ((int) src->address_mode < 0 || src->address_mode >= PL_TEX_ADDRESS_MODE_COUNT) {
        plb_shader_fail(sh, "%s: invalid address mode %d", who, (int) src->address_mode);
        return false;
    }
    o->tex.addr = src->address_mode;
    plb_shader_use_tex(sh, tex);
    if (!require_input_none(sh, resizable ? 0 : o->out_w, resizable ? 0 : o->out_h, who))
        return false;
    if (!tex->params.sampleable) {
        plb_shader_fail(sh, "%s: texture not sampleable!", who);
        return false;
    }
    return true;
}

static int coord_mode(pl_shader sh)
{
    extern int plb_gpu_coord_mode(pl_gpu gpu);
    return plb_gpu_coord_mode(sh->params.gpu);
}

// ---------------------------------------------------------------------------------------
// direct samplers
// ---------------------------------------------------------------------------------------

static bool sample_simple(pl_shader sh, const struct pl_sample_src *src, int linear,
                          const char *who)
{
    struct src_info si;
    if (!setup_src(sh, src, &si, true, linear == 1, who))
        return false;
    sh->src = PLB_SRC_PLANES;
    struct plb_plane_src *pl = &sh->px.planes[0];
    memset(pl, 0, sizeof(*pl));
    pl->kind = PLB_PLANE_DIRECT;
    pl->tex = si.tex;
    pl->rect = si.rect;
    // BEST (sample_direct): linear if the format can
    pl->linear = linear < 0 ? !!(src->tex->params.format->caps & PL_FMT_CAP_LINEAR) : linear;
    pl->scale = si.scale;
    for (int k = 0; k < 4; k++)
        pl->map[k] = k;
    sh->px.nplanes = 1;
    sh->px.coord_mode = coord_mode(sh);
    return true;
}

bool pl_shader_sample_direct(pl_shader sh, const struct pl_sample_src *src)
{
    return sample_simple(sh, src, -1, "pl_shader_sample_direct");
}

bool pl_shader_sample_nearest(pl_shader sh, const struct pl_sample_src *src)
{
    return sample_simple(sh, src, 0, "pl_shader_sample_nearest");
}

bool pl_shader_sample_bilinear(pl_shader sh, const struct pl_sample_src *src)
{
    return sample_simple(sh, src, 1, "pl_shader_sample_bilinear");
}

// Built-in fast scalers (sampling.c:319-478): a handful of bilinear fetches at positions
// computed from the sub-texel phase. All of them need a linearly sampleable texture.
static bool sample_fast(pl_shader sh, const struct pl_sample_src *src, int kind, float threshold,
                        const char *who, const char *step)
{
    struct src_info si;
    if (!setup_src(sh, src, &si, true, true, who))
        return false;
    sh->src = PLB_SRC_PLANES;
    plb_shader_describe(sh, step);
    struct plb_plane_src *pl = &sh->px.planes[0];
    memset(pl, 0, sizeof(*pl));
    pl->kind = kind;
    pl->tex = si.tex;
    pl->rect = si.rect;
    pl->linear = 1;
    pl->scale = si.scale;
    pl->over_rx = si.rx;
    pl->over_ry = si.ry;
    pl->over_threshold = PLB_CLAMP(threshold, 0.0f, 0.5f);
    for (int k = 0; k < 4; k++)
        pl->map[k] = k;
    sh->px.nplanes = 1;
    sh->px.coord_mode = coord_mode(sh);
    return true;
}

bool pl_shader_sample_bicubic(pl_shader sh, const struct pl_sample_src *src)
{
    return sample_fast(sh, src, PLB_PLANE_BICUBIC, 0.0f, "pl_shader_sample_bicubic", "bicubic");
}

bool pl_shader_sample_hermite(pl_shader sh, const struct pl_sample_src *src)
{
    return sample_fast(sh, src, PLB_PLANE_HERMITE, 0.0f, "pl_shader_sample_hermite", "hermite");
}

bool pl_shader_sample_gaussian(pl_shader sh, const struct pl_sample_src *src)
{
    return sample_fast(sh, src, PLB_PLANE_GAUSSIAN, 0.0f, "pl_shader_sample_gaussian", "gaussian");
}

bool pl_shader_sample_oversample(pl_shader sh, const struct pl_sample_src *src, float threshold)
{
    return sample_fast(sh, src, PLB_PLANE_OVERSAMPLE, threshold, "pl_shader_sample_oversample",
                       "oversample");
}

// ---------------------------------------------------------------------------------------
// debanding
// ---------------------------------------------------------------------------------------

const struct pl_deband_params pl_deband_default_params = { PL_DEBAND_DEFAULTS };

void pl_shader_deband(pl_shader sh, const struct pl_sample_src *src,
                      const struct pl_deband_params *params)
{
    struct src_info si;
    if (!setup_src(sh, src, &si, false, false, "pl_shader_deband"))
        return;
    params = params ? params : &pl_deband_default_params;
    sh->src = PLB_SRC_PLANES;
    plb_shader_describe(sh, "debanding");
    struct plb_plane_src *pl = &sh->px.planes[0];
    memset(pl, 0, sizeof(*pl));
    pl->kind = PLB_PLANE_DEBAND;
    pl->tex = si.tex;
    pl->rect = si.rect;
    pl->scale = si.scale;
    for (int k = 0; k < 4; k++)
        pl->map[k] = k;
    pl->comp_mask = si.mask & ~0x8u;
    pl->iterations = params->iterations;
    pl->radius = params->radius;
    // constants exactly as sampling.c:224-226, :261, :266 compute them
    pl->threshold = params->threshold / (1000 * si.scale);
    pl->grain_str = params->grain > 0 ? (float) (params->grain / (1000.0 * si.scale)) : 0.0f;
    memcpy(pl->neutral, params->grain_neutral, sizeof(pl->neutral));
    sh->px.nplanes = 1;
    // the deband shader runs 1:1 on texel centres (NEAREST), coordinates are exact
    sh->px.coord_mode = PLB_COORD_EXACT;
}

// ---------------------------------------------------------------------------------------
// scaler LUT handling shared by polar / ortho
// ---------------------------------------------------------------------------------------

#define SCALER_LUT_SIZE   256
#define SCALER_LUT_CUTOFF 1e-3f

static bool sampler_update(pl_shader sh, struct plb_sampler_obj *obj,
                           const struct pl_filter_config *cfg, bool polar)
{
    pl_gpu gpu = sh->params.gpu;
    if (obj->filter && pl_filter_config_eq(&obj->filter->params.config, cfg))
        return true;
    pl_filter_free(&obj->filter);
    struct pl_filter_params fp = {
        .config = *cfg,
        .lut_entries = SCALER_LUT_SIZE,
        .cutoff = polar ? SCALER_LUT_CUTOFF : 0.0f,
        .max_row_size = polar ? 0 : (int) (gpu->limits.max_tex_2d_dim / 4),
        .row_stride_align = polar ? 0 : 4,
    };
    obj->filter = pl_filter_generate(sh->log, &fp);
    if (!obj->filter) {
        plb_shader_fail(sh, "Failed initializing %s filter!", polar ? "polar" : "separated");
        return false;
    }
    pl_filter f = obj->filter;
    size_t n = polar ? SCALER_LUT_SIZE : (size_t) SCALER_LUT_SIZE * f->row_stride;
    free(obj->host_lut);
    obj->host_lut = malloc(n * sizeof(float));
    if (polar || f->radius != f->radius_zero) {
        memcpy(obj->host_lut, f->weights, n * sizeof(float));
    } else {
        // fill_ortho_lut, sampling.c:919-936: positive-lobe filters use the linear sampling
        // trick, the LUT stores (w0+w1, w1/(w0+w1)) pairs
        for (int i = 0; i < SCALER_LUT_SIZE; i++) {
            const float *w = f->weights + (size_t) i * f->row_stride;
            float *row = obj->host_lut + (size_t) i * f->row_stride;
            int k = 0;
            for (; k < f->row_size; k += 2) {
                const float w0 = w[k], w1 = w[k + 1];
                row[k] = w0 + w1;
                row[k + 1] = w1 / (w0 + w1);
            }
            for (; k < f->row_stride; k++)
                row[k] = k >= 4 ? row[k - 4] : 0;
        }
    }
    if (obj->lut_floats != n) {
        plb_dev_free(gpu, obj->dev_lut);
        obj->dev_lut = plb_dev_alloc(gpu, n * sizeof(float));
        obj->lut_floats = n;
    }
    if (!obj->dev_lut || !plb_dev_upload(gpu, obj->dev_lut, obj->host_lut, n * sizeof(float))) {
        plb_shader_fail(sh, "Failed initializing scaler LUT!");
        return false;
    }
    return true;
}

bool pl_shader_sample_polar(pl_shader sh, const struct pl_sample_src *src,
                            const struct pl_sample_filter_params *params)
{
    if (!params || !params->filter.polar) {
        plb_shader_fail(sh, "Trying to use polar sampling with a non-polar filter?");
        return false;
    }
    struct src_info si;
    if (!setup_src(sh, src, &si, false, false, "pl_shader_sample_polar"))
        return false;
    struct plb_sampler_obj *obj = plb_shader_obj(sh, params->lut, PLB_OBJ_SAMPLER);
    if (!obj)
        return false;

    // anti-alias widening when downscaling, sampling.c:608-616
    float inv_scale = 1.0 / PLB_MIN(si.rx, si.ry);
    inv_scale = PLB_MAX(inv_scale, 1.0);
    if (params->no_widening)
        inv_scale = 1.0;
    struct pl_filter_config cfg = params->filter;
    cfg.antiring = PLB_DEF(cfg.antiring, params->antiring);
    cfg.blur = PLB_DEF(cfg.blur, 1.0f) * inv_scale;
    if (!sampler_update(sh, obj, &cfg, true))
        return false;

    sh->src = PLB_SRC_POLAR;
    plb_shader_describe(sh, "polar scaling");
    struct plb_polar_launch *p = &sh->polar;
    p->src = si.tex;
    p->rect = si.rect;
    p->lut = obj->dev_lut;
    p->host_lut = obj->host_lut;
    p->lut_entries = SCALER_LUT_SIZE;
    p->radius = obj->filter->radius;
    p->radius_zero = obj->filter->radius_zero;
    p->antiring = cfg.antiring;
    p->scale = si.scale;
    p->comp_mask = si.mask;
    p->allow_fast = true;
    sh->px.coord_mode = coord_mode(sh);
    // the reference turns this into a 32x(max_threads/32) compute shader when radius < 6
    // (sampling.c:668-699); we mirror the geometry because peak detection statistics are
    // defined per work group
    if (!params->no_compute && obj->filter->radius < 6.0)
        plb_shader_try_compute(sh, 32, 32, false);
    return true;
}

bool pl_shader_sample_ortho2(pl_shader sh, const struct pl_sample_src *src,
                             const struct pl_sample_filter_params *params)
{
    if (!params || params->filter.polar) {
        plb_shader_fail(sh, "Trying to use separated sampling with a polar filter?");
        return false;
    }
    struct src_info si;
    if (!setup_src(sh, src, &si, false, true, "pl_shader_sample_ortho2"))
        return false;
    int pass;
    if (fabs(si.rx - 1.0f) < 1e-6f) {
        pass = 0; // SEP_VERT
    } else if (fabs(si.ry - 1.0f) < 1e-6f) {
        pass = 1; // SEP_HORIZ
    } else {
        plb_shader_fail(sh, "Trying to use pl_shader_sample_ortho with a pl_sample_src that "
                        "requires scaling in multiple directions (rx=%f, ry=%f), this is not "
                        "possible!", si.rx, si.ry);
        return false;
    }
    struct plb_sampler_obj *obj = plb_shader_obj(sh, params->lut, PLB_OBJ_SAMPLER);
    if (!obj)
        return false;
    if (pass != 0) {
        if (!obj->pass2)
            obj->pass2 = calloc(1, sizeof(*obj->pass2));
        obj = obj->pass2;
    }
    const float ratio = pass ? si.rx : si.ry;
    float inv_scale = 1.0 / ratio;
    inv_scale = PLB_MAX(inv_scale, 1.0);
    if (params->no_widening)
        inv_scale = 1.0;
    struct pl_filter_config cfg = params->filter;
    cfg.antiring = PLB_DEF(cfg.antiring, params->antiring);
    cfg.blur = PLB_DEF(cfg.blur, 1.0f) * inv_scale;
    if (!sampler_update(sh, obj, &cfg, false))
        return false;

    sh->src = PLB_SRC_ORTHO;
    plb_shader_describe(sh, "separable scaling");
    struct plb_ortho_launch *p = &sh->ortho;
    p->src = si.tex;
    p->rect = si.rect;
    p->lut = obj->dev_lut;
    p->lut_entries = SCALER_LUT_SIZE;
    p->row_size = obj->filter->row_size;
    p->row_stride = obj->filter->row_stride;
    p->use_linear = obj->filter->radius == obj->filter->radius_zero;
    p->antiring = cfg.antiring;
    p->ratio = ratio;
    p->scale = si.scale;
    p->comp_mask = si.mask;
    p->pass = pass;
    sh->px.coord_mode = coord_mode(sh);
    return true;
}

// ---------------------------------------------------------------------------------------
// colour ops
// ---------------------------------------------------------------------------------------

// decimal text of printf("%f"), which is what the reference's GLSL carries for these
static double glsl6(double x)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "%f", x);
    return strtod(buf, NULL);
}

static const float PQ_M1 = 2610./4096 * 1./4, PQ_M2 = 2523./4096 * 128, PQ_C1 = 3424./4096,
                   PQ_C2 = 2413./4096 * 32, PQ_C3 = 2392./4096 * 32;
static const float HLG_A = 0.17883277, HLG_B = 0.28466892, HLG_C = 0.55991073;
// Panasonic V-Log / Sony S-Log published curve constants (float-typed, like the reference's)
static const float VLOG_B = 0.00873, VLOG_C = 0.241514, VLOG_D = 0.598206;
static const float SLOG_A = 0.432699, SLOG_B = 0.037584, SLOG_C = 0.616596 + 0.03,
                   SLOG_P = 3.538813, SLOG_Q = 0.030001, SLOG_K2 = 155.0 / 219.0;
static const float HLG_REF = 1000.0 / PL_COLOR_SDR_WHITE;

void plb_pq_literal(struct plb_pq_consts *q)
{
    memset(q, 0, sizeof(*q));
    q->m1 = (float) glsl6(PQ_M1);
    q->m2 = (float) glsl6(PQ_M2);
    q->c1 = (float) glsl6(PQ_C1);
    q->c2 = (float) glsl6(PQ_C2);
    q->c3 = (float) glsl6(PQ_C3);
    q->inv_m1 = (float) (1.0 / glsl6(PQ_M1));
    q->inv_m2 = (float) (1.0 / glsl6(PQ_M2));
    q->to_pq = (float) glsl6(PL_COLOR_SDR_WHITE / 10000);
    q->from_pq = (float) glsl6(10000 / PL_COLOR_SDR_WHITE);
    q->inv_from_pq = (float) (1.0 / glsl6(10000 / PL_COLOR_SDR_WHITE));
    q->hlg_a = (float) glsl6(HLG_A);
    q->hlg_b = (float) glsl6(HLG_B);
    q->hlg_c = (float) glsl6(HLG_C);
    q->hlg_inv_a = (float) (1.0 / glsl6(HLG_A));
    q->inv_2pi = (float) glsl6(0.5f / M_PI);
    q->vlog_b = (float) glsl6(VLOG_B);
    q->vlog_d = (float) glsl6(VLOG_D);
    q->vlog_inv_c = (float) (1.0 / glsl6(VLOG_C));
    q->vlog_c10 = (float) glsl6(VLOG_C / M_LN10);
    q->slog_b = (float) glsl6(SLOG_B);
    q->slog_c = (float) glsl6(SLOG_C);
    q->slog_inv_a = (float) (1.0 / glsl6(SLOG_A));
    q->slog_a10 = (float) glsl6(SLOG_A / M_LN10);
    q->slog_p = (float) glsl6(SLOG_P);
    q->slog_inv_p = (float) (1.0 / glsl6(SLOG_P));
    q->slog_q = (float) glsl6(SLOG_Q);
    q->slog_k2 = (float) glsl6(SLOG_K2);
    q->slog_inv_k2 = (float) (1.0 / glsl6(SLOG_K2));
}

void pl_shader_set_alpha(pl_shader sh, struct pl_color_repr *repr, enum pl_alpha_mode mode)
{
    if (!plb_shader_require_color(sh, "pl_shader_set_alpha"))
        return;
    const bool src_has_alpha = repr->alpha == PL_ALPHA_INDEPENDENT ||
                               repr->alpha == PL_ALPHA_PREMULTIPLIED;
    const bool dst_not_premul = mode == PL_ALPHA_INDEPENDENT || mode == PL_ALPHA_NONE;
    if (repr->alpha == PL_ALPHA_PREMULTIPLIED && dst_not_premul) {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_ALPHA);
        if (!op)
            return;
        op->u.alpha.from = PL_ALPHA_PREMULTIPLIED;
        op->u.alpha.to = PL_ALPHA_INDEPENDENT;
        repr->alpha = PL_ALPHA_INDEPENDENT;
    }
    if (repr->alpha == PL_ALPHA_INDEPENDENT && mode == PL_ALPHA_PREMULTIPLIED) {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_ALPHA);
        if (!op)
            return;
        op->u.alpha.from = PL_ALPHA_INDEPENDENT;
        op->u.alpha.to = PL_ALPHA_PREMULTIPLIED;
        repr->alpha = PL_ALPHA_PREMULTIPLIED;
    }
    if (src_has_alpha && mode == PL_ALPHA_NONE) {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_ALPHA);
        if (!op)
            return;
        op->u.alpha.from = PL_ALPHA_UNKNOWN; // "color.a = 1.0"
        op->u.alpha.to = PL_ALPHA_NONE;
        repr->alpha = PL_ALPHA_NONE;
    }
}

static void push_matrix(pl_shader sh, const pl_transform3x3 *tr)
{
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_MATRIX);
    if (!op)
        return;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            op->u.matrix.m[3*i + j] = tr->mat.m[i][j];
        op->u.matrix.c[i] = tr->c[i];
    }
    op->u.matrix.has_c = 1;
}

// pl_shader_dovi_reshape host part, colorspace.c:118-169: pack the per-piece coefficients and
// the MMR weights the way the GLSL indexes them
static void plb_record_dovi_reshape(pl_shader sh, const struct pl_dovi_metadata *data)
{
    if (sh->px.has_dovi) {
        plb_shader_fail(sh, "only one Dolby Vision reshaping step per pass is supported");
        return;
    }
    struct plb_dovi_tab *d = &sh->px.dovi;
    memset(d, 0, sizeof(*d));
    bool any = false;
    for (int c = 0; c < 3; c++) {
        const struct pl_reshape_data *comp = &data->comp[c];
        d->num_pivots[c] = comp->num_pivots;
        if (!comp->num_pivots)
            continue;
        if (comp->num_pivots < 2 || comp->num_pivots > 9) {
            plb_shader_fail(sh, "invalid number of Dolby Vision pivots: %d", comp->num_pivots);
            return;
        }
        any = true;
        int mmr_idx = 0;
        for (int i = 0; i < comp->num_pivots - 1; i++) {
            float *co = d->coeffs[c][i];
            if (comp->method[i] == 0) {
                for (int k = 0; k < 3; k++)
                    co[k] = comp->poly_coeffs[i][k];
                co[3] = 0.0f; // order 0 signals polynomial
            } else if (comp->method[i] == 1 && comp->mmr_order[i] >= 1 && comp->mmr_order[i] <= 3) {
                co[3] = (float) comp->mmr_order[i];
                co[0] = comp->mmr_constant[i];
                co[1] = (float) mmr_idx;
                for (int j = 0; j < comp->mmr_order[i]; j++) {
                    float *m = &d->mmr[c][mmr_idx][0];
                    m[0] = comp->mmr_coeffs[i][j][0];
                    m[1] = comp->mmr_coeffs[i][j][1];
                    m[2] = comp->mmr_coeffs[i][j][2];
                    m[3] = 0.0f;
                    m[4] = comp->mmr_coeffs[i][j][3];
                    m[5] = comp->mmr_coeffs[i][j][4];
                    m[6] = comp->mmr_coeffs[i][j][5];
                    m[7] = comp->mmr_coeffs[i][j][6];
                    mmr_idx += 2;
                }
            } else {
                plb_shader_fail(sh, "invalid Dolby Vision mapping method");
                return;
            }
        }
        for (int i = 0; i < 7; i++)
            d->pivots[c][i] = i < comp->num_pivots - 2 ? comp->pivots[i + 1] : 1e9f;
        d->lo[c] = comp->pivots[0];
        d->hi[c] = comp->pivots[comp->num_pivots - 1];
    }
    if (!any)
        return;
    if (plb_shader_push_op(sh, PLB_OP_DOVI_RESHAPE))
        sh->px.has_dovi = 1;
}

// Dolby Vision always outputs BT.2020-referred HPE LMS (colorspace.c:392-422)
static void plb_record_dovi_lms(pl_shader sh, const struct pl_dovi_metadata *data)
{
    pl_matrix3x3 lms2rgb = {{
        { 3.06441879, -2.16597676,  0.10155818},
        {-0.65612108,  1.78554118, -0.12943749},
        { 0.01736321, -0.04725154,  1.03004253},
    }};
    pl_matrix3x3_mul(&lms2rgb, &data->linear);
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_DOVI_LMS);
    if (!op)
        return;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            op->u.matrix.m[3 * i + j] = lms2rgb.m[i][j];
    }
}

void pl_shader_dovi_reshape(pl_shader sh, const struct pl_dovi_metadata *data)
{
    if (!plb_shader_require_color(sh, "pl_shader_dovi_reshape") || !data)
        return;
    plb_record_dovi_reshape(sh, data);
}

void pl_shader_decode_color(pl_shader sh, struct pl_color_repr *repr,
                            const struct pl_color_adjustment *params)
{
    if (!plb_shader_require_color(sh, "pl_shader_decode_color"))
        return;
    if (repr->sys == PL_COLOR_SYSTEM_DOLBYVISION) {
        if (!repr->dovi) {
            plb_shader_fail(sh, "pl_shader_decode_color: Dolby Vision without metadata");
            return;
        }
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_SCALE);
        if (!op)
            return;
        op->u.scale.s = pl_color_repr_normalize(repr); // colorspace.c:287 (updates repr)
        pl_shader_dovi_reshape(sh, repr->dovi);
    }

    const enum pl_color_system orig_sys = repr->sys;
    const struct pl_dovi_metadata *dovi = repr->dovi;
    pl_transform3x3 tr = pl_color_repr_decode(repr, params);
    if (memcmp(&tr, &pl_transform3x3_identity, sizeof(tr)))
        push_matrix(sh, &tr);

    switch (orig_sys) {
    case PL_COLOR_SYSTEM_BT_2020_C:
    case PL_COLOR_SYSTEM_BT_2100_PQ:
    case PL_COLOR_SYSTEM_BT_2100_HLG: {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_DECODE_SPECIAL);
        if (op)
            op->u.special.sys = orig_sys;
        break;
    }
    case PL_COLOR_SYSTEM_DOLBYVISION: {
        // PQ EOTF -> (hard-coded HPE LMS->RGB * dovi->linear) -> PQ OETF, colorspace.c:392-422
        plb_record_dovi_lms(sh, dovi);
        break;
    }
    default:
        break;
    }

    if (params && params->gamma == 0) {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_SCALE);
        if (op)
            op->u.scale.s = 0.0f;
    } else if (params && params->gamma != 1) {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_GAMMA);
        if (op)
            op->u.gamma.g = 1 / params->gamma;
    }
    pl_shader_set_alpha(sh, repr, PL_ALPHA_INDEPENDENT);
}

void pl_shader_encode_color(pl_shader sh, const struct pl_color_repr *repr)
{
    if (!plb_shader_require_color(sh, "pl_shader_encode_color"))
        return;
    if (repr->alpha == PL_ALPHA_PREMULTIPLIED) {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_ALPHA);
        if (!op)
            return;
        op->u.alpha.from = PL_ALPHA_INDEPENDENT;
        op->u.alpha.to = PL_ALPHA_PREMULTIPLIED;
    }
    switch (repr->sys) {
    case PL_COLOR_SYSTEM_BT_2020_C:
    case PL_COLOR_SYSTEM_BT_2100_PQ:
    case PL_COLOR_SYSTEM_BT_2100_HLG: {
        struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_ENCODE_SPECIAL);
        if (op)
            op->u.special.sys = repr->sys;
        break;
    }
    case PL_COLOR_SYSTEM_DOLBYVISION:
        plb_shader_fail(sh, "Cannot un-apply dolbyvision yet (no inverse reshaping)!");
        return;
    default:
        break;
    }
    // skip when the representation is plain full-range RGB at native depth (:546-552)
    bool skip = true;
    skip &= PLB_DEF(repr->sys, PL_COLOR_SYSTEM_RGB) == PL_COLOR_SYSTEM_RGB;
    skip &= PLB_DEF(repr->levels, PL_COLOR_LEVELS_FULL) == PL_COLOR_LEVELS_FULL;
    skip &= !repr->bits.sample_depth || !repr->bits.color_depth ||
            repr->bits.sample_depth == repr->bits.color_depth;
    skip &= !repr->bits.bit_shift;
    if (!skip) {
        struct pl_color_repr copy = *repr;
        pl_transform3x3 tr = pl_color_repr_decode(&copy, NULL);
        pl_transform3x3_invert(&tr);
        push_matrix(sh, &tr);
    }
}

void plb_fill_trc(struct plb_trc_consts *t, const struct pl_color_space *csp, bool delin)
{
    memset(t, 0, sizeof(*t));
    t->trc = csp->transfer;
    pl_color_space_nominal_luma_ex(pl_nominal_luma_params(
        .color = csp,
        .metadata = PL_HDR_METADATA_HDR10,
        .scaling = PL_HDR_NORM,
        .out_min = &t->csp_min,
        .out_max = &t->csp_max,
    ));
    const pl_matrix3x3 rgb2xyz = pl_get_rgb2xyz_matrix(pl_raw_primaries_get(csp->primaries));
    for (int i = 0; i < 3; i++)
        t->luma[i] = rgb2xyz.m[1][i];
    switch (csp->transfer) {
    case PL_COLOR_TRC_BT_1886: {
        const float lb = powf(t->csp_min, 1/2.4f);
        const float lw = powf(t->csp_max, 1/2.4f);
        const float a = powf(lw - lb, 2.4f);
        const float b = lb / (lw - lb);
        t->a = delin ? (float) (1.0 / a) : a;
        t->b = b;
        break;
    }
    case PL_COLOR_TRC_HLG: {
        const float y = 1.2f * powf(1.111f, log2f(t->csp_max / HLG_REF));
        const float b = sqrtf(3 * powf(t->csp_min / t->csp_max, 1 / y));
        t->a = y;
        t->b = b;
        break;
    }
    case PL_COLOR_TRC_SCRGB:
        t->a = delin ? (float) glsl6(PL_COLOR_SDR_WHITE / PL_COLOR_SCRGB_WHITE)
                     : (float) glsl6(PL_COLOR_SCRGB_WHITE / PL_COLOR_SDR_WHITE);
        break;
    default:
        break;
    }
    if (delin) {
        t->black_scaled = pl_color_space_is_black_scaled(csp) &&
                          csp->transfer != PL_COLOR_TRC_HLG &&
                          (t->csp_max != 1 || t->csp_min != 0);
    }
}

static bool trc_supported(pl_shader sh, enum pl_color_transfer trc)
{
    if (trc < 0 || trc >= PL_COLOR_TRC_COUNT) {
        plb_shader_fail(sh, "invalid transfer function %d", (int) trc);
        return false;
    }
    return true;
}

void plb_record_linearize(pl_shader sh, const struct pl_color_space *csp)
{
    if (csp->transfer == PL_COLOR_TRC_LINEAR || !trc_supported(sh, csp->transfer))
        return;
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_LINEARIZE);
    if (op)
        plb_fill_trc(&op->u.trc, csp, false);
}

void plb_record_delinearize(pl_shader sh, const struct pl_color_space *csp)
{
    if (csp->transfer == PL_COLOR_TRC_LINEAR || !trc_supported(sh, csp->transfer))
        return;
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_DELINEARIZE);
    if (op)
        plb_fill_trc(&op->u.trc, csp, true);
}

void pl_shader_linearize(pl_shader sh, const struct pl_color_space *csp)
{
    if (plb_shader_require_color(sh, "pl_shader_linearize"))
        plb_record_linearize(sh, csp);
}

void pl_shader_delinearize(pl_shader sh, const struct pl_color_space *csp)
{
    if (plb_shader_require_color(sh, "pl_shader_delinearize"))
        plb_record_delinearize(sh, csp);
}

// Source stage of pl_render_image_mix (renderer.c:3906-3966): `n` cached frames of the output
// size, sampled 1:1, taken to linear light (`lin_csp` != NULL: more than one frame) and summed
// with the normalised weights.
bool plb_shader_mix(pl_shader sh, int n, const pl_tex *texs, const float *weights, int w, int h,
                    const struct pl_color_space *lin_csp)
{
    if (sh->src != PLB_SRC_NONE || sh->px.prog.n) {
        plb_shader_fail(sh, "frame mixing needs an empty shader");
        return false;
    }
    if (n < 1 || n > PLB_MAX_MIX) {
        plb_shader_fail(sh, "frame mixing: %d frames (max %d)", n, PLB_MAX_MIX);
        return false;
    }
    struct plb_mix_launch *m = &sh->mix;
    m->n = n;
    for (int i = 0; i < n; i++) {
        if (texs[i]->params.w != w || texs[i]->params.h != h) {
            plb_shader_fail(sh, "frame mixing: cached frame %d has the wrong size", i);
            return false;
        }
        m->tex[i] = plb_tex_priv(texs[i])->d;
        m->weight[i] = weights[i];
        plb_shader_use_tex(sh, texs[i]);
    }
    m->linearize = lin_csp != NULL;
    if (lin_csp)
        plb_fill_trc(&m->lin, lin_csp, false);
    sh->src = PLB_SRC_MIX;
    sh->out_w = w;
    sh->out_h = h;
    plb_shader_describe(sh, n > 1 ? "frame mixing" : "frame mixing (1 frame)");
    return true;
}

// pl_shader_deinterlace (shaders/deinterlacing.c:32-370): source stage over the interleaved
// frame `src->cur.top` and, for the temporal algorithms, its neighbours.
const struct pl_deinterlace_params pl_deinterlace_default_params = { PL_DEINTERLACE_DEFAULTS };

void pl_shader_deinterlace(pl_shader sh, const struct pl_deinterlace_source *src,
                           const struct pl_deinterlace_params *params)
{
    params = params ? params : &pl_deinterlace_default_params;
    if (!src || !src->cur.top) {
        plb_shader_fail(sh, "pl_shader_deinterlace: no source texture");
        return;
    }
    if (sh->src != PLB_SRC_NONE || sh->px.prog.n) {
        plb_shader_fail(sh, "pl_shader_deinterlace: the shader already has a source");
        return;
    }
    const struct pl_tex_params *tp = &src->cur.top->params;
    if ((sh->out_w && sh->out_w != tp->w) || (sh->out_h && sh->out_h != tp->h)) {
        plb_shader_fail(sh, "pl_shader_deinterlace: incompatible output size");
        return;
    }
    uint8_t mask = src->component_mask ? src->component_mask : 0xFu;
    mask &= (1u << tp->format->num_components) - 1u;
    if (!mask) {
        plb_shader_fail(sh, "pl_shader_deinterlace: empty component mask?");
        return;
    }
    const enum pl_field first_field = src->first_field ? src->first_field : PL_FIELD_TOP;
    bool intra_only = params->algo != PL_DEINTERLACE_YADIF;
    if (params->algo == PL_DEINTERLACE_BWDIF) {
        intra_only = (!src->prev.top && src->field == first_field) ||
                     (!src->next.top && src->field != first_field);
    }
    pl_tex prev = src->cur.top, next = src->cur.top;
    if (!intra_only && src->prev.top)
        prev = src->prev.top;
    if (!intra_only && src->next.top)
        next = src->next.top;
    if (prev->params.w != tp->w || prev->params.h != tp->h || next->params.w != tp->w ||
        next->params.h != tp->h) {
        plb_shader_fail(sh, "pl_shader_deinterlace: reference fields differ in size");
        return;
    }
    struct plb_deint_launch *d = &sh->deint;
    d->cur = plb_tex_priv(src->cur.top)->d;
    d->prev = plb_tex_priv(prev)->d;
    d->next = plb_tex_priv(next)->d;
    d->field = src->field;
    d->prev2_is_prev = src->field == first_field;   // prev2 = field == first ? prev : cur
    d->next2_is_cur = src->field == first_field;    // next2 = field == first ? cur : next
    d->algo = params->algo;
    d->skip_spatial_check = params->skip_spatial_check;
    d->intra_only = intra_only;
    d->comp_mask = mask;
    plb_shader_use_tex(sh, src->cur.top);
    plb_shader_use_tex(sh, prev);
    plb_shader_use_tex(sh, next);
    sh->src = PLB_SRC_DEINT;
    sh->out_w = tp->w;
    sh->out_h = tp->h;
    plb_shader_describe(sh, "deinterlacing");
}

const struct pl_sigmoid_params pl_sigmoid_default_params = { PL_SIGMOID_DEFAULTS };

static void push_sigmoid(pl_shader sh, const struct pl_sigmoid_params *params, int kind)
{
    params = params ? params : &pl_sigmoid_default_params;
    const float center = PLB_DEF(params->center, pl_sigmoid_default_params.center);
    const float slope = PLB_DEF(params->slope, pl_sigmoid_default_params.slope);
    struct plb_color_op *op = plb_shader_push_op(sh, kind);
    if (!op)
        return;
    op->u.sigmoid.center = center;
    op->u.sigmoid.slope = slope;
    op->u.sigmoid.offset = 1.0 / (1 + expf(slope * center));
    op->u.sigmoid.scale = 1.0 / (1 + expf(slope * (center - 1))) - op->u.sigmoid.offset;
}

void pl_shader_sigmoidize(pl_shader sh, const struct pl_sigmoid_params *params)
{
    if (plb_shader_require_color(sh, "pl_shader_sigmoidize"))
        push_sigmoid(sh, params, PLB_OP_SIGMOID);
}

void pl_shader_unsigmoidize(pl_shader sh, const struct pl_sigmoid_params *params)
{
    if (plb_shader_require_color(sh, "pl_shader_unsigmoidize"))
        push_sigmoid(sh, params, PLB_OP_UNSIGMOID);
}

// pl_shader_extract_features (shaders/colorspace.c:1383-1404): the perceptual intensity (the I
// of IPT-PQ) of the colour, the input of the contrast-recovery low-pass. GLSL:
//   lms = <0.020300> * M * color.rgb   -- scalar * mat3 first, then the vector product
void pl_shader_extract_features(pl_shader sh, struct pl_color_space csp)
{
    if (!plb_shader_require_color(sh, "pl_shader_extract_features"))
        return;
    plb_shader_describe(sh, "feature extraction");
    pl_shader_linearize(sh, &csp);
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_FEATURES);
    if (!op)
        return;
    const pl_matrix3x3 m = pl_ipt_rgb2lms(pl_raw_primaries_get(csp.primaries));
    const float k = (float) glsl6(PL_COLOR_SDR_WHITE / 10000);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            op->u.features.m[3 * i + j] = k * (float) m.m[i][j];
    }
}

// ---------------------------------------------------------------------------------------
// custom LUTs (shaders/lut.c:35-280, include/libplacebo/shaders/lut.h)
// ---------------------------------------------------------------------------------------

// 64-bit FNV-1a of the text: only identity matters (the reference hashes with its own function;
// the value is opaque to callers and merely keys the upload)
static uint64_t lut_text_hash(const char *s, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++)
        h = (h ^ (uint8_t) s[i]) * 0x100000001b3ull;
    return h;
}

struct cube_cursor { const char *p, *end; };

static bool cube_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

static void cube_skip_space(struct cube_cursor *c)
{
    while (c->p < c->end && cube_space(*c->p))
        c->p++;
}

// next line without its terminator, stripped; advances past it
static void cube_line(struct cube_cursor *c, const char **beg, const char **end)
{
    const char *b = c->p, *e = b;
    while (e < c->end && *e != '\n')
        e++;
    c->p = e < c->end ? e + 1 : e;
    while (b < e && cube_space(*b))
        b++;
    while (e > b && cube_space(e[-1]))
        e--;
    *beg = b;
    *end = e;
}

// keyword followed by a blank or the end of the line (pl_str_eatstart0)
static bool cube_key(const char **b, const char *e, const char *key)
{
    const size_t n = strlen(key);
    if ((size_t) (e - *b) < n || memcmp(*b, key, n))
        return false;
    *b += n;
    return true;
}

static bool cube_float(const char *b, const char *e, float *out)
{
    char buf[64];
    while (b < e && cube_space(*b))
        b++;
    while (e > b && cube_space(e[-1]))
        e--;
    const size_t n = e - b;
    if (!n || n >= sizeof(buf))
        return false;
    memcpy(buf, b, n);
    buf[n] = 0;
    char *endp;
    *out = strtof(buf, &endp); // correctly rounded decimal -> binary32, as the reference's parser
    return endp == buf + n;
}

static bool cube_int(const char *b, const char *e, int *out)
{
    char buf[32];
    const size_t n = e - b;
    if (!n || n >= sizeof(buf))
        return false;
    memcpy(buf, b, n);
    buf[n] = 0;
    char *endp;
    const long v = strtol(buf, &endp, 10);
    *out = (int) v;
    return endp == buf + n;
}

static bool cube_triple(const char *b, const char *e, float v[3])
{
    for (int i = 0; i < 3; i++) {
        const char *t = b;
        while (t < e && *t != ' ')
            t++;
        if (!cube_float(b, i < 2 ? t : e, &v[i]))
            return false;
        b = t < e ? t + 1 : e;
    }
    return true;
}

struct pl_custom_lut *pl_lut_parse_cube(pl_log log, const char *cstr, size_t cstr_len)
{
    struct pl_custom_lut *lut = calloc(1, sizeof(*lut));
    struct cube_cursor c = { cstr, cstr + cstr_len };
    lut->signature = lut_text_hash(cstr, cstr_len);
    int entries = 0;
    float min[3] = { 0.0, 0.0, 0.0 }, max[3] = { 1.0, 1.0, 1.0 };
    float *data = NULL;

    // header: everything up to the first line that starts with a digit or '-'
    while (c.p < c.end && !((*c.p >= '0' && *c.p <= '9') || *c.p == '-')) {
        const char *b, *e;
        cube_line(&c, &b, &e);
        if (b == e)
            continue;
        if (cube_key(&b, e, "TITLE")) {
            PLB_INFO(log, "Loading LUT: %.*s", (int) (e - b), b);
            continue;
        }
        const bool is3d = cube_key(&b, e, "LUT_3D_SIZE");
        if (is3d || cube_key(&b, e, "LUT_1D_SIZE")) {
            while (b < e && cube_space(*b))
                b++;
            int size;
            if (!cube_int(b, e, &size)) {
                PLB_ERR(log, "Failed parsing dimension '%.*s'", (int) (e - b), b);
                goto error;
            }
            if (size <= 0 || size > (is3d ? 1024 : 65536)) {
                PLB_ERR(log, "Invalid %s size: %d", is3d ? "3DLUT" : "1DLUT", size);
                goto error;
            }
            lut->size[0] = size;
            lut->size[1] = lut->size[2] = is3d ? size : 0;
            entries = is3d ? size * size * size : size;
            continue;
        }
        const bool is_min = cube_key(&b, e, "DOMAIN_MIN");
        if (is_min || cube_key(&b, e, "DOMAIN_MAX")) {
            while (b < e && cube_space(*b))
                b++;
            if (!cube_triple(b, e, is_min ? min : max)) {
                PLB_ERR(log, "Failed parsing domain: '%.*s'", (int) (e - b), b);
                goto error;
            }
            continue;
        }
        if (cube_key(&b, e, "#"))
            continue;
        PLB_WARN(log, "Unhandled .cube line: %.*s", (int) (e - b), b);
    }
    if (!entries) {
        PLB_ERR(log, "Missing LUT size specification?");
        goto error;
    }
    for (int i = 0; i < 3; i++) {
        if (max[i] - min[i] < 1e-6) {
            PLB_ERR(log, "Invalid domain range: [%f, %f]", min[i], max[i]);
            goto error;
        }
    }

    data = malloc(sizeof(float[3]) * entries);
    lut->data = data;
    for (int n = 0; n < entries; n++) {
        for (int ch = 0; ch < 3; ch++) {
            const char *b = c.p;
            while (c.p < c.end && strchr("0123456789.-+e", *c.p) && *c.p)
                c.p++;
            float num;
            if (b == c.p) {
                if (c.p == c.end)
                    PLB_ERR(log, "Failed parsing LUT: Unexpected EOF, expected %d entries, got %d",
                            entries * 3, n * 3 + ch + 1);
                else
                    PLB_ERR(log, "Failed parsing LUT: Unexpected '%c', expected digit", *c.p);
                goto error;
            }
            if (!cube_float(b, c.p, &num)) {
                PLB_ERR(log, "Failed parsing float value '%.*s'", (int) (c.p - b), b);
                goto error;
            }
            data[3 * n + ch] = (num - min[ch]) / (max[ch] - min[ch]); // rescale to 0.0 - 1.0
            cube_skip_space(&c);
        }
    }
    cube_skip_space(&c);
    if (c.p < c.end)
        PLB_WARN(log, "Extra data after LUT?... ignoring '%c'", *c.p);
    return lut;

error:
    free(data);
    free(lut);
    return NULL;
}

void pl_lut_free(struct pl_custom_lut **lut)
{
    if (!lut || !*lut)
        return;
    free((void *) (*lut)->data);
    free(*lut);
    *lut = NULL;
}

// pl_shader_custom_lut (shaders/lut.c:212-280): optional shaper matrix, the LUT (1-D: each channel
// through its own component with linear interpolation; 3-D: tetrahedral interpolation between
// four texels, lut.c:773-809), optional output matrix. The table is uploaded once per signature
// as 4 floats per entry (the reference's rgba32f texture, fill_lut :196-210).
void pl_shader_custom_lut(pl_shader sh, const struct pl_custom_lut *lut, pl_shader_obj *lut_state)
{
    if (!lut)
        return;
    int dims;
    if (lut->size[0] > 0 && lut->size[1] > 0 && lut->size[2] > 0) {
        dims = 3;
    } else if (lut->size[0] > 0 && !lut->size[1] && !lut->size[2]) {
        dims = 1;
    } else {
        plb_shader_fail(sh, "Invalid dimensions %dx%dx%d for pl_custom_lut, must be 1D or 3D!",
                        lut->size[0], lut->size[1], lut->size[2]);
        return;
    }
    if (!plb_shader_require_color(sh, "pl_shader_custom_lut"))
        return;
    struct plb_lut_obj *obj = plb_shader_obj(sh, lut_state, PLB_OBJ_LUT);
    if (!obj)
        return;
    pl_gpu gpu = sh->params.gpu;
    const size_t entries = (size_t) lut->size[0] * PLB_DEF(lut->size[1], 1) * PLB_DEF(lut->size[2], 1);
    if (!obj->dev_lut || obj->signature != lut->signature || memcmp(obj->size, lut->size, sizeof(obj->size))) {
        float *tmp = malloc(entries * 4 * sizeof(float));
        for (size_t i = 0; i < entries; i++) {
            tmp[4 * i + 0] = lut->data[3 * i + 0];
            tmp[4 * i + 1] = lut->data[3 * i + 1];
            tmp[4 * i + 2] = lut->data[3 * i + 2];
            tmp[4 * i + 3] = 0.0f;
        }
        plb_dev_free(gpu, obj->dev_lut);
        obj->dev_lut = plb_dev_alloc(gpu, entries * 4 * sizeof(float));
        const bool ok = obj->dev_lut && plb_dev_upload(gpu, obj->dev_lut, tmp, entries * 4 * sizeof(float));
        free(tmp);
        if (!ok) {
            plb_shader_fail(sh, "pl_shader_custom_lut: failed generating LUT object");
            return;
        }
        obj->signature = lut->signature;
        memcpy(obj->size, lut->size, sizeof(obj->size));
    }
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_CUSTOM_LUT);
    if (!op)
        return;
    plb_shader_describe(sh, dims == 3 ? "custom 3DLUT" : "custom 1DLUT");
    op->u.custom_lut.lut = obj->dev_lut;
    memcpy(op->u.custom_lut.size, lut->size, sizeof(lut->size));
    static const pl_matrix3x3 zero = {0};
    op->u.custom_lut.has_in = memcmp(&lut->shaper_in, &zero, sizeof(zero)) != 0;
    op->u.custom_lut.has_out = memcmp(&lut->shaper_out, &zero, sizeof(zero)) != 0;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            op->u.custom_lut.shaper_in[3 * i + j] = lut->shaper_in.m[i][j];
            op->u.custom_lut.shaper_out[3 * i + j] = lut->shaper_out.m[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------
// dithering (dithering.c:109-274)
// ---------------------------------------------------------------------------------------

const struct pl_dither_params pl_dither_default_params = { PL_DITHER_DEFAULTS };

static float approx_gamma(enum pl_color_transfer trc)
{
    switch (trc) {
    case PL_COLOR_TRC_UNKNOWN:
    case PL_COLOR_TRC_LINEAR:
    case PL_COLOR_TRC_SCRGB:     return 1.0f;
    case PL_COLOR_TRC_PRO_PHOTO:
    case PL_COLOR_TRC_GAMMA18:   return 1.8f;
    case PL_COLOR_TRC_GAMMA20:   return 2.0f;
    case PL_COLOR_TRC_GAMMA24:   return 2.4f;
    case PL_COLOR_TRC_GAMMA26:
    case PL_COLOR_TRC_ST428:     return 2.6f;
    case PL_COLOR_TRC_GAMMA28:   return 2.8f;
    case PL_COLOR_TRC_SRGB:
    case PL_COLOR_TRC_BT_1886:
    case PL_COLOR_TRC_GAMMA22:   return 2.2f;
    default:                     return 2.0f; // PQ, HLG, camera logs
    }
}

void pl_shader_dither(pl_shader sh, int new_depth, pl_shader_obj *dither_state,
                      const struct pl_dither_params *params)
{
    if (!plb_shader_require_color(sh, "pl_shader_dither"))
        return;
    if (new_depth <= 0 || new_depth > 256) {
        PLB_WARN(sh->log, "Invalid dither depth: %d.. ignoring", new_depth);
        return;
    }
    params = params ? params : &pl_dither_default_params;
    if (params->lut_size < 0 || params->lut_size > 8) {
        plb_shader_fail(sh, "Invalid `lut_size` specified: %d", params->lut_size);
        return;
    }
    if (new_depth >= 64) {
        // (1 << depth) - 1 is not representable; the reference's GLSL would overflow too
        PLB_WARN(sh->log, "dither depth %d too large.. ignoring", new_depth);
        return;
    }

    enum pl_dither_method method = params->method;
    const bool is_lut = method == PL_DITHER_BLUE_NOISE || method == PL_DITHER_ORDERED_LUT;
    struct plb_dither_obj *obj = NULL;
    int lut_size = 0;
    if (is_lut) {
        if (!dither_state) {
            PLB_WARN(sh->log, "LUT-based dither method specified but no dither state object "
                     "given, falling back to non-LUT based methods.");
            method = PL_DITHER_ORDERED_FIXED;
        } else {
            obj = plb_shader_obj(sh, dither_state, PLB_OBJ_DITHER);
            if (!obj)
                return;
            lut_size = 1 << PLB_DEF(params->lut_size, pl_dither_default_params.lut_size);
            if (!obj->dev_lut || obj->lut_size != lut_size || obj->method != (int) method) {
                pl_gpu gpu = sh->params.gpu;
                float *tmp = malloc(sizeof(float) * lut_size * lut_size);
                if (method == PL_DITHER_ORDERED_LUT)
                    pl_generate_bayer_matrix(tmp, lut_size);
                else
                    pl_generate_blue_noise(tmp, lut_size);
                plb_dev_free(gpu, obj->dev_lut);
                obj->dev_lut = plb_dev_alloc(gpu, sizeof(float) * lut_size * lut_size);
                bool ok = obj->dev_lut &&
                          plb_dev_upload(gpu, obj->dev_lut, tmp, sizeof(float) * lut_size * lut_size);
                free(tmp);
                if (!ok) {
                    plb_shader_fail(sh, "failed uploading dither matrix");
                    return;
                }
                obj->lut_size = lut_size;
                obj->method = method;
            }
        }
    }

    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_DITHER);
    if (!op)
        return;
    struct plb_op_dither *d = &op->u.dither;
    d->method = method;
    d->lut = obj ? obj->dev_lut : NULL;
    d->lut_size = lut_size;
    d->new_depth = new_depth;
    d->temporal = params->temporal;
    d->frame_index = sh->params.index;
    d->gamma = approx_gamma(params->transfer);
    d->rot[0] = 1; d->rot[1] = 0; d->rot[2] = 0; d->rot[3] = 1;
    if (params->temporal) {
        // dithering.c:185-199; mat2 built from row-major floats, uploaded column-major
        const int phase = sh->params.index % 8;
        const float r = phase * (M_PI / 2);
        const float m = phase < 4 ? 1 : -1;
        d->rot[0] = cos(r);
        d->rot[1] = -sin(r);
        d->rot[2] = sin(r) * m;
        d->rot[3] = cos(r) * m;
    }
    d->scale = (float) ((1ull << new_depth) - 1);
    d->inv_scale = 1.0f / d->scale;
    d->fast_pow2 = (method == PL_DITHER_BLUE_NOISE || method == PL_DITHER_ORDERED_LUT) &&
                   !params->temporal && lut_size > 0 && !(lut_size & (lut_size - 1)) &&
                   !(d->gamma != 1.0f && new_depth <= 4);
}
