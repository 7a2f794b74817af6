/*
 * TEST INFRASTRUCTURE — not part of the product, never linked into libplacebo_b200.so.
 *
 * Parameter harvester: a thin flat-C shim over the UNMODIFIED reference library
 * (oracle/_ref/libplacebo_ref.so, built by oracle/build_ref.sh from /root/reference).
 * It exposes two kinds of entry points to the Python tests (ctypes):
 *
 *   hvt_*  direct calls into the reference's CPU parameter derivation ("T0"):
 *          pl_filter_generate (src/filters.c:186), pl_color_repr_decode
 *          (src/colorspace.c:1717), pl_tone_map_generate (src/tone_mapping.c:147),
 *          pl_gamut_map_generate (src/gamut_mapping.c:412), pl_generate_blue_noise
 *          (src/dither.c:175), ...
 *   hvs_*  run a real pl_shader_* generator on the reference's dummy pl_gpu
 *          (src/dummy.c) and dump everything it binds: LUT textures
 *          (pl_tex_dummy_data), SSBO contents, uniform variables, specialization
 *          constants, compute work-group geometry and the GLSL text (the GLSL is for human
 *          review of the restatement in oracle/pl_oracle.c; it cannot execute here).
 *
 * Compiled against the reference headers in place (-I/root/reference/src/include etc.,
 * flags recorded by build_ref.sh in _ref/cflags.txt); nothing from the reference is copied
 * into this repository.
 */
#include "common.h"
#include "gpu.h"
#include "shaders.h"

#include <libplacebo/dummy.h>
#include <libplacebo/filters.h>
#include <libplacebo/tone_mapping.h>
#include <libplacebo/gamut_mapping.h>
#include <libplacebo/dither.h>
#include <libplacebo/colorspace.h>
#include <libplacebo/shaders/sampling.h>
#include <libplacebo/shaders/colorspace.h>
#include <libplacebo/shaders/dithering.h>
#include <libplacebo/shaders/film_grain.h>
#include <libplacebo/shaders/lut.h>
#include <libplacebo/renderer.h>
#include <libplacebo/utils/upload.h>

#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

// Fixtures of the reference's own tests (av1_grain_data, h274_grain_data, dovi_meta)
#include "tests/utils.h"

#define HV_API __attribute__((visibility("default")))

static pl_log hv_log;
static pl_gpu hv_gpu;

static void log_cb(void *priv, enum pl_log_level level, const char *msg)
{
    if (level <= PL_LOG_WARN)
        fprintf(stderr, "[ref] %s\n", msg);
}

static void hv_init(void)
{
    if (hv_gpu)
        return;
    hv_log = pl_log_create(PL_API_VER, pl_log_params(
        .log_cb = log_cb,
        .log_level = PL_LOG_WARN,
    ));
    hv_gpu = pl_gpu_dummy_create(hv_log, NULL);
}

// ---------------------------------------------------------------------------------------
// hvt_*: direct T0 calls. The wrappers only use the PUBLIC libplacebo API, so the very same
// source (t0_flat.inc) is compiled a second time against include/libplacebo_b200.h by
// tests/c/shim.c (prefix plt_*), which lets the tests compare product and reference call by
// call.
// ---------------------------------------------------------------------------------------
#define T0_FN(n) hvt_##n
#define T0_API HV_API
#define T0_INIT() hv_init()
#define T0_LOG hv_log
#define T0_GPU hv_gpu
#include "t0_flat.inc"

// Raw bytes of the reference test fixtures, so that the golden generator can store them
HV_API int hvt_fixture(const char *name, void *out, int cap)
{
    const void *p = NULL;
    int n = 0;
    if (!strcmp(name, "av1_grain_data")) { p = &av1_grain_data; n = sizeof(av1_grain_data); }
    else if (!strcmp(name, "dovi_meta")) { p = &dovi_meta; n = sizeof(dovi_meta); }
    if (!p || n > cap)
        return -1;
    memcpy(out, p, n);
    return n;
}

// ---------------------------------------------------------------------------------------
// hvs_*: run real shader generators on the dummy GPU and dump what they bind
// ---------------------------------------------------------------------------------------

static void dump_res(const char *dir, const struct pl_shader_res *res)
{
    char path[1024];
    mkdir(dir, 0755);
    snprintf(path, sizeof(path), "%s/glsl.txt", dir);
    FILE *f = fopen(path, "w");
    fputs(res->glsl, f);
    fclose(f);

    snprintf(path, sizeof(path), "%s/meta.txt", dir);
    f = fopen(path, "w");
    fprintf(f, "name %s\ninput %d\noutput %d\ngroup %d %d\nshmem %zu\n", res->name,
            res->input, res->output, res->compute_group_size[0],
            res->compute_group_size[1], res->compute_shmem);
    for (int i = 0; i < res->info->num_steps; i++)
        fprintf(f, "step %s\n", res->info->steps[i]);

    for (int i = 0; i < res->num_constants; i++) {
        const struct pl_shader_const *c = &res->constants[i];
        switch (c->type) {
        case PL_VAR_SINT:  fprintf(f, "const %s int %d\n", c->name, *(const int *) c->data); break;
        case PL_VAR_UINT:  fprintf(f, "const %s uint %u\n", c->name, *(const unsigned *) c->data); break;
        case PL_VAR_FLOAT: fprintf(f, "const %s float %.9g\n", c->name, *(const float *) c->data); break;
        default: break;
        }
    }

    for (int i = 0; i < res->num_variables; i++) {
        const struct pl_shader_var *v = &res->variables[i];
        int n = v->var.dim_v * v->var.dim_m * v->var.dim_a;
        fprintf(f, "var %s type %d dims %d %d %d dyn %d :", v->var.name, v->var.type,
                v->var.dim_v, v->var.dim_m, v->var.dim_a, v->dynamic);
        for (int k = 0; k < n; k++) {
            if (v->var.type == PL_VAR_FLOAT)
                fprintf(f, " %.9g", ((const float *) v->data)[k]);
            else if (v->var.type == PL_VAR_SINT)
                fprintf(f, " %d", ((const int *) v->data)[k]);
            else
                fprintf(f, " %u", ((const unsigned *) v->data)[k]);
        }
        fprintf(f, "\n");
    }

    for (int i = 0; i < res->num_vertex_attribs; i++) {
        const struct pl_shader_va *va = &res->vertex_attribs[i];
        fprintf(f, "va %s :", va->attr.name);
        for (int k = 0; k < 4; k++) {
            const float *p = va->data[k];
            fprintf(f, " %.9g %.9g", p[0], p[1]);
        }
        fprintf(f, "\n");
    }

    for (int i = 0; i < res->num_descriptors; i++) {
        const struct pl_shader_desc *sd = &res->descriptors[i];
        if (sd->desc.type == PL_DESC_SAMPLED_TEX || sd->desc.type == PL_DESC_STORAGE_IMG) {
            pl_tex tex = sd->binding.object;
            const void *data = pl_tex_dummy_data(tex);
            size_t size = (size_t) tex->params.w * PL_DEF(tex->params.h, 1) *
                          PL_DEF(tex->params.d, 1) * tex->params.format->texel_size;
            fprintf(f, "desc %d %s tex %s %d %d %d sample %d addr %d data %d\n", i,
                    sd->desc.name, tex->params.format->name, tex->params.w,
                    tex->params.h, tex->params.d, sd->binding.sample_mode,
                    sd->binding.address_mode, data ? 1 : 0);
            if (data) {
                snprintf(path, sizeof(path), "%s/desc_%d.bin", dir, i);
                FILE *b = fopen(path, "wb");
                fwrite(data, 1, size, b);
                fclose(b);
            }
        } else {
            pl_buf buf = sd->binding.object;
            fprintf(f, "desc %d %s buf %zu\n", i, sd->desc.name, buf->params.size);
            const void *data = pl_buf_dummy_data(buf);
            if (data) {
                snprintf(path, sizeof(path), "%s/desc_%d.bin", dir, i);
                FILE *b = fopen(path, "wb");
                fwrite(data, 1, buf->params.size, b);
                fclose(b);
            }
        }
    }
    fclose(f);
}

static pl_tex placeholder(const char *fmt, int w, int h)
{
    hv_init();
    return pl_tex_dummy_create(hv_gpu, pl_tex_dummy_params(
        .w = w,
        .h = h,
        .format = pl_find_named_fmt(hv_gpu, fmt),
    ));
}

static const struct pl_shader_res *finish(pl_shader sh, const char *dir)
{
    if (pl_shader_is_failed(sh))
        return NULL;
    const struct pl_shader_res *res = pl_shader_finalize(sh);
    if (res)
        dump_res(dir, res);
    return res;
}

// pl_shader_sample_polar / pl_shader_sample_ortho2 (pass selected by the generator)
HV_API int hvs_sampler(const char *dir, const char *filter, int polar, const char *fmt,
                       int sw, int sh_, int dw, int dh, int comps, float antiring)
{
    const struct pl_filter_config *c = pl_find_filter_config(filter, 0);
    if (!c)
        return -1;
    pl_tex tex = placeholder(fmt, sw, sh_);
    struct pl_sample_src src = {
        .tex = tex,
        .new_w = dw,
        .new_h = dh,
        .components = comps,
    };
    pl_shader_obj lut = NULL;
    struct pl_sample_filter_params fp = {
        .filter = *c,
        .lut = &lut,
        .antiring = antiring,
    };
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu ));
    bool ok = polar ? pl_shader_sample_polar(sh, &src, &fp)
                    : pl_shader_sample_ortho2(sh, &src, &fp);
    int ret = ok && finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_shader_obj_destroy(&lut);
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}

HV_API int hvs_deband(const char *dir, const char *fmt, int w, int h, int comps,
                      int iterations, float threshold, float radius, float grain,
                      const float *neutral3, float scale, int index)
{
    pl_tex tex = placeholder(fmt, w, h);
    struct pl_sample_src src = { .tex = tex, .components = comps, .scale = scale };
    struct pl_deband_params dp = {
        .iterations = iterations,
        .threshold = threshold,
        .radius = radius,
        .grain = grain,
    };
    memcpy(dp.grain_neutral, neutral3, sizeof(dp.grain_neutral));
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu, .index = index ));
    pl_shader_deband(sh, &src, &dp);
    int ret = finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}

// preset: 0 = pl_color_map_default_params, 1 = pl_color_map_high_quality_params
// tone_fn / gamut_fn optional overrides (NULL keeps the preset)
HV_API int hvs_color_map(const char *dir, const float *src_csp, const float *dst_csp, int preset,
                         const char *tone_fn, const char *gamut_fn, int with_state)
{
    hv_init();
    struct pl_color_map_params par = preset ? pl_color_map_high_quality_params
                                            : pl_color_map_default_params;
    if (tone_fn)
        par.tone_mapping_function = pl_find_tone_map_function(tone_fn);
    if (gamut_fn)
        par.gamut_mapping = pl_find_gamut_map_function(gamut_fn);
    par.contrast_recovery = 0; // feature map is out of scope
    pl_shader_obj state = NULL;
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu ));
    // emulate an upstream sampling step so that the shader has a color signature
    pl_tex tex = placeholder("rgba16hf", 64, 64);
    pl_shader_sample_direct(sh, pl_sample_src( .tex = tex ));
    pl_shader_color_map_ex(sh, &par, pl_color_map_args(
        .src = csp_from_flat(src_csp),
        .dst = csp_from_flat(dst_csp),
        .state = with_state ? &state : NULL,
    ));
    int ret = finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_shader_obj_destroy(&state);
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}

HV_API int hvs_decode_color(const char *dir, int sys, int levels, int sample_depth,
                            int color_depth, int bit_shift, int use_dovi)
{
    hv_init();
    struct pl_color_repr repr = {
        .sys = sys,
        .levels = levels,
        .bits = { sample_depth, color_depth, bit_shift },
        .dovi = use_dovi ? &dovi_meta : NULL,
    };
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu ));
    pl_tex tex = placeholder("rgba16hf", 64, 64);
    pl_shader_sample_direct(sh, pl_sample_src( .tex = tex ));
    pl_shader_decode_color(sh, &repr, NULL);
    int ret = finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}

HV_API int hvs_dither(const char *dir, int method, int lut_size, int depth, int temporal,
                      int transfer, int index)
{
    hv_init();
    struct pl_dither_params dp = {
        .method = method,
        .lut_size = lut_size,
        .temporal = temporal,
        .transfer = transfer,
    };
    pl_shader_obj state = NULL;
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu, .index = index ));
    pl_tex tex = placeholder("rgba16hf", 64, 64);
    pl_shader_sample_direct(sh, pl_sample_src( .tex = tex ));
    pl_shader_dither(sh, depth, &state, &dp);
    int ret = finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_shader_obj_destroy(&state);
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}

HV_API int hvs_detect_peak(const char *dir, const float *csp, int hq)
{
    hv_init();
    pl_shader_obj state = NULL;
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu ));
    pl_tex tex = placeholder("rgba16hf", 64, 64);
    pl_shader_sample_direct(sh, pl_sample_src( .tex = tex ));
    bool ok = pl_shader_detect_peak(sh, csp_from_flat(csp), &state,
                                    hq ? &pl_peak_detect_high_quality_params
                                       : &pl_peak_detect_default_params);
    int ret = ok && finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_shader_obj_destroy(&state);
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}

// AV1 film grain on one plane, using the reference test fixture (tests/utils.h:240-262)
// plane: 0 luma (r16 w x h), 1 = merged chroma (rg16, w/2 x h/2 with luma tex)
HV_API int hvs_av1_grain(const char *dir, int w, int h, int plane, unsigned seed, int overlap,
                         int bit_depth, int levels, int sys)
{
    hv_init();
    struct pl_film_grain_data data = {
        .type = PL_FILM_GRAIN_AV1,
        .seed = seed,
        .params.av1 = av1_grain_data,
    };
    data.params.av1.overlap = overlap;
    struct pl_color_repr repr = {
        .sys = sys,
        .levels = levels,
        .bits = { 16, bit_depth, 0 },
    };
    pl_tex luma = placeholder("r16", w, h);
    pl_tex tex = plane ? placeholder("rg16", w / 2, h / 2) : luma;
    struct pl_film_grain_params par = {
        .data = data,
        .tex = tex,
        .luma_tex = luma,
        .repr = &repr,
        .components = plane ? 2 : 1,
        .component_mapping = { plane ? PL_CHANNEL_CB : PL_CHANNEL_Y, PL_CHANNEL_CR },
        .luma_comp = 0,
    };
    pl_shader_obj state = NULL;
    pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu ));
    bool ok = pl_shader_film_grain(sh, &state, &par);
    int ret = ok && finish(sh, dir) ? 0 : -2;
    pl_shader_free(&sh);
    pl_shader_obj_destroy(&state);
    if (tex != luma)
        pl_tex_destroy(hv_gpu, &tex);
    pl_tex_destroy(hv_gpu, &luma);
    return ret;
}

// H.274 film grain on one plane with the reference test fixture (tests/utils.h:263-277):
// `fmt` r16 / rg16 / rgba16 plane of `comps` components mapped to channels map[0..comps).
// The dump holds the grain data base (desc 0) and the uniforms (seed, scale_factor, bounds,
// comp_model_value) the shader was given.
HV_API int hvs_h274_grain(const char *dir, const char *fmt, int w, int h, int comps,
                          const int *map, unsigned seed, int bit_depth, int levels, int sys)
{
    hv_init();
    struct pl_film_grain_data data = {
        .type = PL_FILM_GRAIN_H274,
        .seed = seed,
        .params.h274 = h274_grain_data,
    };
    struct pl_color_repr repr = {
        .sys = sys,
        .levels = levels,
        .bits = { 16, bit_depth, 0 },
    };
    pl_tex tex = placeholder(fmt, w, h);
    struct pl_film_grain_params par = {
        .data = data,
        .tex = tex,
        .repr = &repr,
        .components = comps,
    };
    for (int i = 0; i < comps; i++)
        par.component_mapping[i] = map[i];
    int ret = -3;
    if (pl_needs_film_grain(&par)) {
        pl_shader_obj state = NULL;
        pl_shader sh = pl_shader_alloc(hv_log, pl_shader_params( .gpu = hv_gpu ));
        bool ok = pl_shader_film_grain(sh, &state, &par);
        ret = ok && finish(sh, dir) ? 0 : -2;
        pl_shader_free(&sh);
        pl_shader_obj_destroy(&state);
    }
    pl_tex_destroy(hv_gpu, &tex);
    return ret;
}
