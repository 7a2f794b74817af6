// Private layout of pl_shader / pl_shader_obj / pl_dispatch on the CUDA backend.
//
// Reference counterparts: struct pl_shader_t (src/shaders.h:85-119) holds GLSL string
// builders; here it holds a typed description of the same pass: one SOURCE stage (what the
// first pl_shader_sample_* / deband / film_grain call set up, possibly several merged planes)
// followed by a colour-op program. pl_dispatch_finish lowers it to kernel launches
// (csrc/dispatch.cpp), which is where src/dispatch.c:1199-1368 would have generated, compiled
// and run a GLSL pass.
#ifndef PLB_SHADER_PRIV_H_
#define PLB_SHADER_PRIV_H_

#include "internal.h"

#ifdef __cplusplus
extern "C" {
#endif

enum plb_obj_type {
    PLB_OBJ_INVALID = 0,
    PLB_OBJ_SAMPLER,    // PL_SHADER_OBJ_SAMPLER: filter + LUT
    PLB_OBJ_COLOR_MAP,  // tone LUT, gamut 3DLUT, peak detection state
    PLB_OBJ_DITHER,
    PLB_OBJ_GRAIN,
    PLB_OBJ_LUT,        // PL_SHADER_OBJ_LUT: custom 1-D / 3-D LUT
};

struct plb_sampler_obj {
    pl_filter filter;
    float *dev_lut;        // device copy of the weights (polar: 256 floats; ortho: rows)
    float *host_lut;       // host copy in the layout the kernels use
    size_t lut_floats;
    struct plb_sampler_obj *pass2; // second 1-D pass of pl_shader_sample_ortho2
};

struct plb_peak_state {
    struct pl_peak_detect_params params;
    uint32_t *dev_buf;     // pending struct peak_buf_data on the device (3264 B)
    bool pending;
    float avg_pq, max_pq;  // smoothed values (colorspace.c:981-987)
};

struct plb_color_map_obj {
    struct pl_tone_map_params tone_params;
    bool have_tone;
    float *dev_tone_lut;
    size_t tone_lut_size;
    uint64_t gamut_sig;
    bool have_gamut;
    float *dev_gamut_lut;      // rgba32f texels = unorm16 values / 65535
    size_t gamut_texels;
    struct plb_peak_state peak;
};

struct plb_dither_obj {
    float *dev_lut;
    int lut_size;
    int method;
};

struct plb_grain_obj;      // k_grain / film_grain.cpp

struct plb_lut_obj {
    float *dev_lut;
    uint64_t signature;
    int size[3];
};

struct pl_shader_obj_t {
    enum plb_obj_type type;
    pl_gpu gpu;
    int refcount;
    union {
        struct plb_sampler_obj sampler;
        struct plb_color_map_obj cmap;
        struct plb_dither_obj dither;
        struct plb_grain_obj *grain;
        struct plb_lut_obj lut;
    } u;
};

void *plb_shader_obj(pl_shader sh, pl_shader_obj *ptr, enum plb_obj_type type);

enum plb_src_kind {
    PLB_SRC_NONE = 0,
    PLB_SRC_PLANES,     // direct / deband / grain sampling of up to 4 planes
    PLB_SRC_POLAR,
    PLB_SRC_ORTHO,
    PLB_SRC_ERRDIFF,    // pl_shader_error_diffusion (compute-only shader)
    PLB_SRC_MIX,        // frame mixing (pl_render_image_mix)
    PLB_SRC_DEINT,      // pl_shader_deinterlace
};

struct pl_shader_t {
    pl_log log;
    struct pl_shader_params params;
    bool failed;
    bool is_compute;
    int out_w, out_h;          // required output size, 0 = resizable (sh_require)
    int wg_w, wg_h;            // compute work group size requested so far
    bool wg_flexible;
    enum plb_src_kind src;
    struct plb_polar_launch polar;
    struct plb_ortho_launch ortho;
    struct plb_errdiff_launch errdiff;
    struct plb_mix_launch mix;
    struct plb_deint_launch deint;
    struct plb_pixel_launch px;    // planes + program + (later) target
    // shader objects referenced by this shader (kept alive until dispatch)
    pl_shader_obj objs[16];
    int num_objs;
    pl_dispatch dp;            // owner, if created through pl_dispatch_begin
    // friendly names of the recorded operations (sh_describe, src/shaders.h:155-165)
    const char *steps[24];
    int num_steps;
    // textures read by the recorded pass (ordering against the copy-engine streams)
    pl_tex texs[12];
    int num_texs;
};
void plb_shader_use_tex(pl_shader sh, pl_tex tex);

// `desc` must have static storage duration
void plb_shader_describe(pl_shader sh, const char *desc);

void plb_shader_fail(pl_shader sh, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
struct plb_color_op *plb_shader_push_op(pl_shader sh, int kind);
bool plb_shader_require_color(pl_shader sh, const char *who);
bool plb_shader_try_compute(pl_shader sh, int bw, int bh, bool flexible);

// Per-pass bookkeeping behind pl_dispatch_callback (dispatch.c:65-100 `struct pass`): passes
// are identified by a signature of their kernel kind, steps and geometry; each keeps a ring of
// in-flight cudaEvent timers and the last execution times.
#define PLB_PASS_TIMERS 64
struct plb_pass {
    uint64_t signature;
    struct pl_shader_info_t info;
    const char *steps[24];
    char description[256];
    pl_timer timers[PLB_PASS_TIMERS];
    int timer_head, timer_tail;      // ring of pending timers
    struct pl_dispatch_info stats;
    int sample_idx;
};

struct pl_dispatch_t {
    pl_log log;
    pl_gpu gpu;
    uint8_t current_index;     // frame counter (dispatch.c:1613-1622)
    struct plb_pass *passes[64];
    int num_passes;
    void *info_priv;
    void (*info_callback)(void *priv, const struct pl_dispatch_info *info);
};

// renderer-internal: merge a sub-shader's planes into `sh` (sh_subpass + "color[i] = tmp[j]")
bool plb_shader_merge_plane(pl_shader sh, pl_shader sub, const int8_t map[4], float scale);

void plb_timer_begin(pl_gpu gpu, pl_timer t);
void plb_timer_end(pl_gpu gpu, pl_timer t);
void plb_timer_wait(pl_gpu gpu, pl_timer t);

#ifdef __cplusplus
}
#endif
#endif
