/*
 * libplacebo_b200 — C ABI of the B200-native libplacebo hot path.
 *
 * One header, C11/C++ clean, no CUDA or torch types. It declares the subset of libplacebo's
 * public API (v7.365) that sits on the per-frame path pl_render_image -> pl_shader_* ->
 * pl_dispatch_finish, with the SAME names, argument meaning, struct field names/order and
 * error behaviour (bool/NULL + log, never abort), so that a caller written against
 * <libplacebo/...> recompiles against this library unchanged (include/libplacebo/ holds
 * forwarding stubs). Each block cites the reference header it replaces, relative to
 * /root/reference/src/include/libplacebo/.
 *
 * What is different behind the API: there is no GLSL. `pl_shader_*` calls RECORD typed ops
 * into the pl_shader; pl_dispatch_finish() lowers the op list to hand-written sm_100a
 * kernels (libplacebo_b200/csrc/k_*.cu). pl_shader_finalize() therefore does not exist.
 * The `pl_gpu` comes from pl_cuda_create(), the sibling of pl_vulkan_create /
 * pl_gpu_dummy_create (reference: vulkan.h, dummy.h:36-100).
 */
#ifndef LIBPLACEBO_B200_H_
#define LIBPLACEBO_B200_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#define PL_MAJOR_VER 7
#define PL_API_VER 365
#define PL_B200 1

#if defined(__cplusplus)
# define PL_API_BEGIN extern "C" {
# define PL_API_END }
#else
# define PL_API_BEGIN
# define PL_API_END
#endif
#define PL_API __attribute__((visibility("default")))

PL_API_BEGIN

/* ---- common.h:26-115 -------------------------------------------------------------------- */
typedef struct pl_rect2d { int x0, y0, x1, y1; } pl_rect2d;
typedef struct pl_rect3d { int x0, y0, z0, x1, y1, z1; } pl_rect3d;
typedef struct pl_rect2df { float x0, y0, x1, y1; } pl_rect2df;
typedef struct pl_rect3df { float x0, y0, z0, x1, y1, z1; } pl_rect3df;
#define pl_rect_w(r) ((r).x1 - (r).x0)
#define pl_rect_h(r) ((r).y1 - (r).y0)
#define pl_rect_d(r) ((r).z1 - (r).z0)
#define pl_rect2d_eq(a, b) \
    ((a).x0 == (b).x0 && (a).x1 == (b).x1 && (a).y0 == (b).y0 && (a).y1 == (b).y1)

typedef struct pl_matrix3x3 { float m[3][3]; } pl_matrix3x3;
typedef struct pl_transform3x3 { pl_matrix3x3 mat; float c[3]; } pl_transform3x3;
PL_API extern const pl_matrix3x3 pl_matrix3x3_identity;
PL_API extern const pl_transform3x3 pl_transform3x3_identity;
PL_API void pl_matrix3x3_apply(const pl_matrix3x3 *mat, float vec[3]);
PL_API void pl_matrix3x3_scale(pl_matrix3x3 *mat, float scale);
PL_API void pl_matrix3x3_invert(pl_matrix3x3 *mat);
PL_API void pl_matrix3x3_mul(pl_matrix3x3 *a, const pl_matrix3x3 *b);
PL_API void pl_matrix3x3_rmul(const pl_matrix3x3 *a, pl_matrix3x3 *b);
PL_API void pl_transform3x3_apply(const pl_transform3x3 *t, float vec[3]);
PL_API void pl_transform3x3_scale(pl_transform3x3 *t, float scale);
PL_API void pl_transform3x3_invert(pl_transform3x3 *t);
typedef int pl_rotation;
enum { PL_ROTATION_0 = 0, PL_ROTATION_90 = 1, PL_ROTATION_180 = 2, PL_ROTATION_270 = 3, PL_ROTATION_360 = 4 };

/* ---- log.h:26-81 ------------------------------------------------------------------------ */
enum pl_log_level {
    PL_LOG_NONE = 0, PL_LOG_FATAL, PL_LOG_ERR, PL_LOG_WARN, PL_LOG_INFO, PL_LOG_DEBUG,
    PL_LOG_TRACE, PL_LOG_ALL = PL_LOG_TRACE,
};
struct pl_log_params {
    void (*log_cb)(void *log_priv, enum pl_log_level level, const char *msg);
    void *log_priv;
    enum pl_log_level log_level;
};
#define pl_log_params(...) (&(struct pl_log_params) { __VA_ARGS__ })
typedef const struct pl_log_t { struct pl_log_params params; } *pl_log;
#define pl_log_create pl_log_create_365
PL_API pl_log pl_log_create(int api_ver, const struct pl_log_params *params);
PL_API void pl_log_destroy(pl_log *log);
PL_API void pl_log_simple(void *stream, enum pl_log_level level, const char *msg);
// log.h:88-96 — replace the parameters / the level of a live log; return the previous ones
PL_API struct pl_log_params pl_log_update(pl_log log, const struct pl_log_params *params);
PL_API enum pl_log_level pl_log_level_update(pl_log log, enum pl_log_level level);
// log.h:105 — the same with ANSI colours per level
PL_API void pl_log_color(void *stream, enum pl_log_level level, const char *msg);

/* ---- colorspace.h ----------------------------------------------------------------------- */
enum pl_color_system { /* colorspace.h:30-60 */
    PL_COLOR_SYSTEM_UNKNOWN = 0, PL_COLOR_SYSTEM_BT_601, PL_COLOR_SYSTEM_BT_709,
    PL_COLOR_SYSTEM_SMPTE_240M, PL_COLOR_SYSTEM_BT_2020_NC, PL_COLOR_SYSTEM_BT_2020_C,
    PL_COLOR_SYSTEM_BT_2100_PQ, PL_COLOR_SYSTEM_BT_2100_HLG, PL_COLOR_SYSTEM_DOLBYVISION,
    PL_COLOR_SYSTEM_YCGCO, PL_COLOR_SYSTEM_YCGCO_RE, PL_COLOR_SYSTEM_YCGCO_RO,
    PL_COLOR_SYSTEM_RGB, PL_COLOR_SYSTEM_XYZ, PL_COLOR_SYSTEM_COUNT
};
PL_API bool pl_color_system_is_ycbcr_like(enum pl_color_system sys);
// colorspace.h:54-78 — names and the resolution heuristic
PL_API bool pl_color_system_is_linear(enum pl_color_system sys);
enum pl_channel {
    PL_CHANNEL_NONE = -1, PL_CHANNEL_A = 3, PL_CHANNEL_R = 0, PL_CHANNEL_G = 1,
    PL_CHANNEL_B = 2, PL_CHANNEL_Y = 0, PL_CHANNEL_CB = 1, PL_CHANNEL_CR = 2,
    PL_CHANNEL_U = 1, PL_CHANNEL_V = 2
};
enum pl_color_levels {
    PL_COLOR_LEVELS_UNKNOWN = 0, PL_COLOR_LEVELS_LIMITED, PL_COLOR_LEVELS_FULL,
    PL_COLOR_LEVELS_COUNT,
    PL_COLOR_LEVELS_TV = PL_COLOR_LEVELS_LIMITED, PL_COLOR_LEVELS_PC = PL_COLOR_LEVELS_FULL,
};
enum pl_alpha_mode {
    PL_ALPHA_UNKNOWN = 0, PL_ALPHA_INDEPENDENT, PL_ALPHA_PREMULTIPLIED, PL_ALPHA_NONE,
    PL_ALPHA_MODE_COUNT,
};
struct pl_bit_encoding { int sample_depth, color_depth, bit_shift; };
struct pl_dovi_metadata { /* colorspace.h:145-166 */
    float nonlinear_offset[3];
    pl_matrix3x3 nonlinear;
    pl_matrix3x3 linear;
    struct pl_reshape_data {
        uint8_t num_pivots;
        float pivots[9];
        uint8_t method[8];
        float poly_coeffs[8][3];
        uint8_t mmr_order[8];
        float mmr_constant[8];
        float mmr_coeffs[8][3][7];
    } comp[3];
};
struct pl_color_repr {
    enum pl_color_system sys;
    enum pl_color_levels levels;
    enum pl_alpha_mode alpha;
    struct pl_bit_encoding bits;
    const struct pl_dovi_metadata *dovi;
};
PL_API bool pl_color_repr_equal(const struct pl_color_repr *c1, const struct pl_color_repr *c2);
// colorspace.h:176-215 — representation presets, merging (fields set in `orig` win)
PL_API float pl_color_repr_normalize(struct pl_color_repr *repr);
PL_API enum pl_color_levels pl_color_levels_guess(const struct pl_color_repr *repr);

enum pl_color_primaries { /* colorspace.h:218-244 */
    PL_COLOR_PRIM_UNKNOWN = 0, PL_COLOR_PRIM_BT_601_525, PL_COLOR_PRIM_BT_601_625,
    PL_COLOR_PRIM_BT_709, PL_COLOR_PRIM_BT_470M, PL_COLOR_PRIM_EBU_3213,
    PL_COLOR_PRIM_BT_2020, PL_COLOR_PRIM_APPLE, PL_COLOR_PRIM_ADOBE, PL_COLOR_PRIM_PRO_PHOTO,
    PL_COLOR_PRIM_CIE_1931, PL_COLOR_PRIM_DCI_P3, PL_COLOR_PRIM_DISPLAY_P3,
    PL_COLOR_PRIM_V_GAMUT, PL_COLOR_PRIM_S_GAMUT, PL_COLOR_PRIM_FILM_C,
    PL_COLOR_PRIM_ACES_AP0, PL_COLOR_PRIM_ACES_AP1, PL_COLOR_PRIM_COUNT
};
enum pl_color_transfer { /* colorspace.h:255-280 */
    PL_COLOR_TRC_UNKNOWN = 0, PL_COLOR_TRC_BT_1886, PL_COLOR_TRC_SRGB, PL_COLOR_TRC_LINEAR,
    PL_COLOR_TRC_GAMMA18, PL_COLOR_TRC_GAMMA20, PL_COLOR_TRC_GAMMA22, PL_COLOR_TRC_GAMMA24,
    PL_COLOR_TRC_GAMMA26, PL_COLOR_TRC_GAMMA28, PL_COLOR_TRC_PRO_PHOTO, PL_COLOR_TRC_ST428,
    PL_COLOR_TRC_PQ, PL_COLOR_TRC_HLG, PL_COLOR_TRC_V_LOG, PL_COLOR_TRC_S_LOG1,
    PL_COLOR_TRC_S_LOG2, PL_COLOR_TRC_SCRGB, PL_COLOR_TRC_COUNT
};
PL_API float pl_color_transfer_nominal_peak(enum pl_color_transfer trc);
static inline bool pl_color_transfer_is_hdr(enum pl_color_transfer trc)
{
    return pl_color_transfer_nominal_peak(trc) > 1.0;
}
#define PL_COLOR_SDR_WHITE 203.0f
#define PL_COLOR_SCRGB_WHITE 80.0f
#define PL_COLOR_SDR_CONTRAST 1000.0f
#define PL_COLOR_HDR_BLACK 1e-6f
#define PL_COLOR_HLG_PEAK 1000.0f

struct pl_cie_xy { float x, y; };
struct pl_raw_primaries { struct pl_cie_xy red, green, blue, white; };
PL_API bool pl_raw_primaries_equal(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b);
PL_API bool pl_raw_primaries_similar(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b);
PL_API void pl_raw_primaries_merge(struct pl_raw_primaries *orig, const struct pl_raw_primaries *update);
PL_API const struct pl_raw_primaries *pl_raw_primaries_get(enum pl_color_primaries prim);
// colorspace.h:228-258, 262-266 — names, wide-gamut predicate, resolution heuristic

enum pl_hdr_scaling { PL_HDR_NORM = 0, PL_HDR_SQRT, PL_HDR_NITS, PL_HDR_PQ, PL_HDR_SCALING_COUNT };
PL_API float pl_hdr_rescale(enum pl_hdr_scaling from, enum pl_hdr_scaling to, float x);
enum pl_hdr_metadata_type {
    PL_HDR_METADATA_ANY = 0, PL_HDR_METADATA_NONE, PL_HDR_METADATA_HDR10,
    PL_HDR_METADATA_HDR10PLUS, PL_HDR_METADATA_CIE_Y, PL_HDR_METADATA_TYPE_COUNT,
};
struct pl_hdr_bezier {
    float target_luma;
    float knee_x, knee_y;
    float anchors[15];
    uint8_t num_anchors;
};
struct pl_hdr_metadata { /* colorspace.h:398-432 */
    struct pl_raw_primaries prim;
    float min_luma, max_luma;
    float max_cll;
    float max_fall;
    float scene_max[3];
    float scene_avg;
    struct pl_hdr_bezier ootf;
    float max_pq_y;
    float avg_pq_y;
};
PL_API extern const struct pl_hdr_metadata pl_hdr_metadata_empty;
PL_API extern const struct pl_hdr_metadata pl_hdr_metadata_hdr10;
PL_API bool pl_hdr_metadata_equal(const struct pl_hdr_metadata *a, const struct pl_hdr_metadata *b);
PL_API void pl_hdr_metadata_merge(struct pl_hdr_metadata *orig, const struct pl_hdr_metadata *update);
PL_API bool pl_hdr_metadata_contains(const struct pl_hdr_metadata *data, enum pl_hdr_metadata_type type);

enum pl_rendering_intent {
    PL_INTENT_AUTO = -1, PL_INTENT_PERCEPTUAL = 0, PL_INTENT_RELATIVE_COLORIMETRIC = 1,
    PL_INTENT_SATURATION = 2, PL_INTENT_ABSOLUTE_COLORIMETRIC = 3
};
struct pl_color_space {
    enum pl_color_primaries primaries;
    enum pl_color_transfer transfer;
    struct pl_hdr_metadata hdr;
};
#define pl_color_space(...) (&(struct pl_color_space) { __VA_ARGS__ })
PL_API bool pl_color_space_is_hdr(const struct pl_color_space *csp);
PL_API bool pl_color_space_is_black_scaled(const struct pl_color_space *csp);
// colorspace.h:566-574 — CPU transfer functions (in place, three channels; the same curves the
// pixel kernels evaluate, here with libm like the reference)
struct pl_nominal_luma_params {
    const struct pl_color_space *color;
    enum pl_hdr_metadata_type metadata;
    enum pl_hdr_scaling scaling;
    float *out_min;
    float *out_max;
    float *out_avg;
};
#define pl_nominal_luma_params(...) (&(struct pl_nominal_luma_params) { __VA_ARGS__ })
PL_API void pl_color_space_nominal_luma_ex(const struct pl_nominal_luma_params *params);
PL_API void pl_color_space_merge(struct pl_color_space *orig, const struct pl_color_space *update);
PL_API bool pl_color_space_equal(const struct pl_color_space *c1, const struct pl_color_space *c2);
PL_API void pl_color_space_infer(struct pl_color_space *space);
PL_API void pl_color_space_infer_ref(struct pl_color_space *space, const struct pl_color_space *ref);
PL_API void pl_color_space_infer_map(struct pl_color_space *src, struct pl_color_space *dst);
PL_API extern const struct pl_color_space pl_color_space_unknown;
PL_API extern const struct pl_color_space pl_color_space_srgb;
PL_API extern const struct pl_color_space pl_color_space_bt709;
PL_API extern const struct pl_color_space pl_color_space_hdr10;
PL_API extern const struct pl_color_space pl_color_space_bt2020_hlg;
PL_API extern const struct pl_color_space pl_color_space_monitor;

struct pl_color_adjustment { float brightness, contrast, saturation, hue, gamma, temperature; };
#define PL_COLOR_ADJUSTMENT_NEUTRAL .contrast = 1.0, .saturation = 1.0, .gamma = 1.0,
PL_API extern const struct pl_color_adjustment pl_color_adjustment_neutral;

enum pl_chroma_location {
    PL_CHROMA_UNKNOWN = 0, PL_CHROMA_LEFT, PL_CHROMA_CENTER, PL_CHROMA_TOP_LEFT,
    PL_CHROMA_TOP_CENTER, PL_CHROMA_BOTTOM_LEFT, PL_CHROMA_BOTTOM_CENTER, PL_CHROMA_COUNT,
};
PL_API void pl_chroma_location_offset(enum pl_chroma_location loc, float *x, float *y);
PL_API pl_matrix3x3 pl_get_rgb2xyz_matrix(const struct pl_raw_primaries *prim);
PL_API pl_matrix3x3 pl_get_xyz2rgb_matrix(const struct pl_raw_primaries *prim);

// colorspace.h:665-709 — colour-vision deficiency model (cone distortion). pl_render_params.
// cone_params is outside the CUDA hot path (rejected); only the parameter type is declared.
enum pl_cone {
    PL_CONE_L = 1 << 0, PL_CONE_M = 1 << 1, PL_CONE_S = 1 << 2,
    PL_CONE_NONE = 0,
    PL_CONE_LM = PL_CONE_L | PL_CONE_M, PL_CONE_MS = PL_CONE_M | PL_CONE_S,
    PL_CONE_LS = PL_CONE_L | PL_CONE_S, PL_CONE_LMS = PL_CONE_L | PL_CONE_M | PL_CONE_S,
};
struct pl_cone_params {
    enum pl_cone cones; // cones affected by the vision model
    float strength;     // 0 = complete loss, 1 = unaffected
};
#define pl_cone_params(...) (&(struct pl_cone_params) { __VA_ARGS__ })

PL_API pl_matrix3x3 pl_get_color_mapping_matrix(const struct pl_raw_primaries *src,
                                                const struct pl_raw_primaries *dst,
                                                enum pl_rendering_intent intent);
PL_API pl_matrix3x3 pl_get_adaptation_matrix(struct pl_cie_xy src, struct pl_cie_xy dst);
PL_API bool pl_primaries_superset(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b);
PL_API bool pl_primaries_valid(const struct pl_raw_primaries *prim);
PL_API bool pl_primaries_compatible(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b);
PL_API struct pl_raw_primaries pl_primaries_clip(const struct pl_raw_primaries *src,
                                                 const struct pl_raw_primaries *dst);
PL_API pl_matrix3x3 pl_ipt_rgb2lms(const struct pl_raw_primaries *prim);
PL_API pl_matrix3x3 pl_ipt_lms2rgb(const struct pl_raw_primaries *prim);
PL_API extern const pl_matrix3x3 pl_ipt_lms2ipt;
PL_API extern const pl_matrix3x3 pl_ipt_ipt2lms;
PL_API pl_transform3x3 pl_color_repr_decode(struct pl_color_repr *repr,
                                            const struct pl_color_adjustment *params);

/* ---- filters.h -------------------------------------------------------------------------- */
#define PL_FILTER_MAX_PARAMS 2
struct pl_filter_ctx { float radius; float params[PL_FILTER_MAX_PARAMS]; };
struct pl_filter_function { /* filters.h:38-73 */
    const char *name;
    float radius;
    bool resizable;
    bool tunable[PL_FILTER_MAX_PARAMS];
    float params[PL_FILTER_MAX_PARAMS];
    double (*weight)(const struct pl_filter_ctx *f, double x);
    bool opaque;
};
PL_API extern const struct pl_filter_function pl_filter_function_box;
PL_API extern const struct pl_filter_function pl_filter_function_triangle;
PL_API extern const struct pl_filter_function pl_filter_function_cosine;
PL_API extern const struct pl_filter_function pl_filter_function_hann;
PL_API extern const struct pl_filter_function pl_filter_function_hamming;
PL_API extern const struct pl_filter_function pl_filter_function_welch;
PL_API extern const struct pl_filter_function pl_filter_function_kaiser;
PL_API extern const struct pl_filter_function pl_filter_function_blackman;
PL_API extern const struct pl_filter_function pl_filter_function_bohman;
PL_API extern const struct pl_filter_function pl_filter_function_gaussian;
PL_API extern const struct pl_filter_function pl_filter_function_quadratic;
PL_API extern const struct pl_filter_function pl_filter_function_sinc;
PL_API extern const struct pl_filter_function pl_filter_function_jinc;
PL_API extern const struct pl_filter_function pl_filter_function_sphinx;
PL_API extern const struct pl_filter_function pl_filter_function_cubic;
PL_API extern const struct pl_filter_function pl_filter_function_hermite;
PL_API extern const struct pl_filter_function pl_filter_function_spline16;
PL_API extern const struct pl_filter_function pl_filter_function_spline36;
PL_API extern const struct pl_filter_function pl_filter_function_spline64;
PL_API extern const struct pl_filter_function pl_filter_function_oversample;
PL_API const struct pl_filter_function *pl_find_filter_function(const char *name);
// filters.h:166-167 — NULL-terminated list of all kernels / windows
PL_API extern const struct pl_filter_function * const pl_filter_functions[];
PL_API extern const int pl_num_filter_functions;

enum pl_filter_usage {
    PL_FILTER_UPSCALING = (1 << 0), PL_FILTER_DOWNSCALING = (1 << 1),
    PL_FILTER_FRAME_MIXING = (1 << 2),
    PL_FILTER_SCALING = PL_FILTER_UPSCALING | PL_FILTER_DOWNSCALING,
    PL_FILTER_ALL = PL_FILTER_SCALING | PL_FILTER_FRAME_MIXING,
};
struct pl_filter_config { /* filters.h:163-224 */
    const char *name;
    const char *description;
    enum pl_filter_usage allowed;
    enum pl_filter_usage recommended;
    const struct pl_filter_function *kernel;
    const struct pl_filter_function *window;
    float radius;
    float params[PL_FILTER_MAX_PARAMS];
    float wparams[PL_FILTER_MAX_PARAMS];
    float clamp;
    float blur;
    float taper;
    bool polar;
    float antiring;
};
PL_API bool pl_filter_config_eq(const struct pl_filter_config *a, const struct pl_filter_config *b);
PL_API double pl_filter_sample(const struct pl_filter_config *c, double x);
PL_API extern const struct pl_filter_config pl_filter_spline16;
PL_API extern const struct pl_filter_config pl_filter_spline36;
PL_API extern const struct pl_filter_config pl_filter_spline64;
PL_API extern const struct pl_filter_config pl_filter_nearest;
PL_API extern const struct pl_filter_config pl_filter_box;
PL_API extern const struct pl_filter_config pl_filter_bilinear;
PL_API extern const struct pl_filter_config pl_filter_gaussian;
PL_API extern const struct pl_filter_config pl_filter_oversample;
PL_API extern const struct pl_filter_config pl_filter_sinc;
PL_API extern const struct pl_filter_config pl_filter_lanczos;
PL_API extern const struct pl_filter_config pl_filter_ginseng;
PL_API extern const struct pl_filter_config pl_filter_ewa_jinc;
PL_API extern const struct pl_filter_config pl_filter_ewa_lanczos;
PL_API extern const struct pl_filter_config pl_filter_ewa_lanczossharp;
PL_API extern const struct pl_filter_config pl_filter_ewa_lanczos4sharpest;
PL_API extern const struct pl_filter_config pl_filter_ewa_ginseng;
PL_API extern const struct pl_filter_config pl_filter_ewa_hann;
PL_API extern const struct pl_filter_config pl_filter_bicubic;
PL_API extern const struct pl_filter_config pl_filter_hermite;
PL_API extern const struct pl_filter_config pl_filter_catmull_rom;
PL_API extern const struct pl_filter_config pl_filter_mitchell;
PL_API extern const struct pl_filter_config pl_filter_mitchell_clamp;
PL_API extern const struct pl_filter_config pl_filter_robidoux;
PL_API extern const struct pl_filter_config pl_filter_robidouxsharp;
PL_API extern const struct pl_filter_config pl_filter_ewa_robidoux;
PL_API extern const struct pl_filter_config pl_filter_ewa_robidouxsharp;
PL_API extern const struct pl_filter_config * const pl_filter_configs[];
PL_API extern const int pl_num_filter_configs;
// filters.h:316-329 — name -> configuration list of the older selection API ("none" = built-in
// sampling, aliases have no description); terminated by {0}
struct pl_filter_preset {
    const char *name;
    const struct pl_filter_config *filter;
    const char *description;
};
PL_API extern const struct pl_filter_preset pl_filter_presets[];
PL_API extern const int pl_num_filter_presets;
PL_API const struct pl_filter_preset *pl_find_filter_preset(const char *name);
PL_API const struct pl_filter_config *pl_find_filter_config(const char *name, enum pl_filter_usage usage);

struct pl_filter_params { /* filters.h:310-340 */
    struct pl_filter_config config;
    int lut_entries;
    float cutoff;
    int max_row_size;
    int row_stride_align;
    float filter_scale;     /* deprecated in the reference (no effect); kept for layout */
};
#define pl_filter_params(...) (&(struct pl_filter_params) { __VA_ARGS__ })
typedef const struct pl_filter_t {
    struct pl_filter_params params;
    float radius;
    float radius_zero;
    const float *weights;
    int row_size;
    bool insufficient;
    int row_stride;
} *pl_filter;
PL_API pl_filter pl_filter_generate(pl_log log, const struct pl_filter_params *params);
PL_API void pl_filter_free(pl_filter *filter);

/* ---- tone_mapping.h --------------------------------------------------------------------- */
struct pl_tone_map_params;
struct pl_tone_map_function { /* tone_mapping.h:33-68 */
    const char *name;
    const char *description;
    enum pl_hdr_scaling scaling;
    void (*map)(float *lut, const struct pl_tone_map_params *params);
    void (*map_inverse)(float *lut, const struct pl_tone_map_params *params);
    void *priv;
};
struct pl_tone_map_constants { /* tone_mapping.h:70-114 */
    float knee_adaptation, knee_minimum, knee_maximum, knee_default, knee_offset;
    float slope_tuning, slope_offset;
    float spline_contrast;
    float reinhard_contrast;
    float linear_knee;
    float exposure;
};
#define PL_TONE_MAP_CONSTANTS \
    .knee_adaptation = 0.4f, .knee_minimum = 0.1f, .knee_maximum = 0.8f, \
    .knee_default = 0.4f, .knee_offset = 1.0f, .slope_tuning = 1.5f, .slope_offset = 0.2f, \
    .spline_contrast = 0.5f, .reinhard_contrast = 0.5f, .linear_knee = 0.3f, .exposure = 1.0f,
struct pl_tone_map_params { /* tone_mapping.h:129-167 */
    const struct pl_tone_map_function *function;
    struct pl_tone_map_constants constants;
    enum pl_hdr_scaling input_scaling;
    enum pl_hdr_scaling output_scaling;
    size_t lut_size;
    float input_min, input_max, input_avg;
    float output_min, output_max;
    struct pl_hdr_metadata hdr;
    float param;            /* deprecated in the reference (see constants); kept for layout */
};
PL_API bool pl_tone_map_params_equal(const struct pl_tone_map_params *a, const struct pl_tone_map_params *b);
PL_API void pl_tone_map_params_infer(struct pl_tone_map_params *params);
PL_API bool pl_tone_map_params_noop(const struct pl_tone_map_params *params);
PL_API void pl_tone_map_generate(float *out, const struct pl_tone_map_params *params);
PL_API float pl_tone_map_sample(float x, const struct pl_tone_map_params *params);
PL_API extern const struct pl_tone_map_function pl_tone_map_clip;
PL_API extern const struct pl_tone_map_function pl_tone_map_st2094_40;
PL_API extern const struct pl_tone_map_function pl_tone_map_st2094_10;
PL_API extern const struct pl_tone_map_function pl_tone_map_bt2390;
PL_API extern const struct pl_tone_map_function pl_tone_map_bt2446a;
PL_API extern const struct pl_tone_map_function pl_tone_map_spline;
PL_API extern const struct pl_tone_map_function pl_tone_map_reinhard;
PL_API extern const struct pl_tone_map_function pl_tone_map_mobius;
PL_API extern const struct pl_tone_map_function pl_tone_map_hable;
PL_API extern const struct pl_tone_map_function pl_tone_map_gamma;
PL_API extern const struct pl_tone_map_function pl_tone_map_linear;
PL_API extern const struct pl_tone_map_function pl_tone_map_linear_light;
PL_API extern const struct pl_tone_map_function * const pl_tone_map_functions[];
PL_API extern const int pl_num_tone_map_functions;
PL_API const struct pl_tone_map_function *pl_find_tone_map_function(const char *name);
#define pl_tone_map_auto pl_tone_map_spline

/* ---- gamut_mapping.h -------------------------------------------------------------------- */
struct pl_gamut_map_params;
struct pl_gamut_map_function { /* gamut_mapping.h:31-46 */
    const char *name;
    const char *description;
    void (*map)(float *lut, const struct pl_gamut_map_params *params);
    bool bidirectional;
    void *priv;
};
struct pl_gamut_map_constants {
    float perceptual_deadzone, perceptual_strength, colorimetric_gamma;
    float softclip_knee, softclip_desat;
};
#define PL_GAMUT_MAP_CONSTANTS \
    .colorimetric_gamma = 1.80f, .softclip_knee = 0.70f, .softclip_desat = 0.35f, \
    .perceptual_deadzone = 0.30f, .perceptual_strength = 0.80f,
struct pl_gamut_map_params { /* gamut_mapping.h:68-101 */
    const struct pl_gamut_map_function *function;
    struct pl_raw_primaries input_gamut;
    struct pl_raw_primaries output_gamut;
    float min_luma, max_luma;
    struct pl_gamut_map_constants constants;
    int lut_size_I, lut_size_C, lut_size_h;
    int lut_stride;
};
PL_API bool pl_gamut_map_params_equal(const struct pl_gamut_map_params *a, const struct pl_gamut_map_params *b);
PL_API bool pl_gamut_map_params_noop(const struct pl_gamut_map_params *params);
PL_API void pl_gamut_map_generate(float *out, const struct pl_gamut_map_params *params);
PL_API void pl_gamut_map_sample(float x[3], const struct pl_gamut_map_params *params);
PL_API extern const struct pl_gamut_map_function pl_gamut_map_clip;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_perceptual;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_softclip;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_relative;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_saturation;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_absolute;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_desaturate;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_darken;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_highlight;
PL_API extern const struct pl_gamut_map_function pl_gamut_map_linear;
PL_API extern const struct pl_gamut_map_function * const pl_gamut_map_functions[];
PL_API extern const int pl_num_gamut_map_functions;
PL_API const struct pl_gamut_map_function *pl_find_gamut_map_function(const char *name);

/* ---- dither.h --------------------------------------------------------------------------- */
PL_API void pl_generate_bayer_matrix(float *data, int size);
PL_API void pl_generate_blue_noise(float *data, int size);
#define PL_EDF_MIN_DX (-2)
#define PL_EDF_MAX_DX  (2)
#define PL_EDF_MAX_DY  (2)
struct pl_error_diffusion_kernel {
    const char *name;
    const char *description;
    int shift;
    int pattern[PL_EDF_MAX_DY + 1][PL_EDF_MAX_DX - PL_EDF_MIN_DX + 1];
    int divisor;
};
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_simple;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_false_fs;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_sierra_lite;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_floyd_steinberg;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_atkinson;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_jarvis_judice_ninke;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_stucki;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_burkes;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_sierra2;
PL_API extern const struct pl_error_diffusion_kernel pl_error_diffusion_sierra3;
PL_API extern const struct pl_error_diffusion_kernel * const pl_error_diffusion_kernels[];
PL_API extern const int pl_num_error_diffusion_kernels;
PL_API const struct pl_error_diffusion_kernel *pl_find_error_diffusion_kernel(const char *name);

/* ---- gpu.h ------------------------------------------------------------------------------ */
typedef const char *pl_debug_tag;
struct pl_glsl_version { /* gpu.h:55-79 — kept so callers can probe compute limits */
    int version;
    bool gles, vulkan, compute;
    size_t max_shmem_size;
    uint32_t max_group_threads;
    uint32_t max_group_size[3];
    uint32_t subgroup_size;
    int16_t min_gather_offset, max_gather_offset;
};
struct pl_gpu_limits { /* gpu.h:93-147, member for member (callers read it through pl_gpu) */
    bool thread_safe;
    bool callbacks;
    size_t max_buf_size;
    size_t max_ubo_size;
    size_t max_ssbo_size;
    size_t max_vbo_size;
    size_t max_mapped_size;
    uint64_t max_buffer_texels;
    bool host_cached;
    bool host_ptr_slow;
    size_t max_mapped_vram;
    size_t align_host_ptr;
    uint32_t max_tex_1d_dim, max_tex_2d_dim, max_tex_3d_dim;
    bool blittable_1d_3d;
    bool buf_transfer;
    size_t align_tex_xfer_pitch;
    size_t align_tex_xfer_offset;
    size_t max_variable_comps;
    size_t max_constants;
    bool array_size_constants;
    size_t max_pushc_size;
    size_t align_vertex_stride;
    uint32_t max_dispatch[3];
    uint32_t fragment_queues;
    uint32_t compute_queues;
};
typedef uint64_t pl_handle_caps;
struct pl_gpu_handle_caps { pl_handle_caps tex, buf, sync; };    /* gpu.h:166-170: none here */
struct pl_gpu_pci_address { uint32_t domain, bus, device, function; };
typedef const struct pl_fmt_t *pl_fmt;
typedef const struct pl_gpu_t { /* gpu.h:213-235 */
    pl_log log;
    struct pl_glsl_version glsl;
    struct pl_gpu_limits limits;
    struct pl_gpu_handle_caps export_caps;
    struct pl_gpu_handle_caps import_caps;
    uint8_t uuid[16];
    pl_fmt *formats;
    int num_formats;
    struct pl_gpu_pci_address pci;
} *pl_gpu;

enum pl_fmt_type {
    PL_FMT_UNKNOWN = 0, PL_FMT_UNORM, PL_FMT_SNORM, PL_FMT_UINT, PL_FMT_SINT, PL_FMT_FLOAT,
    PL_FMT_TYPE_COUNT,
};
enum pl_fmt_caps {
    PL_FMT_CAP_SAMPLEABLE = 1 << 0, PL_FMT_CAP_STORABLE = 1 << 1, PL_FMT_CAP_LINEAR = 1 << 2,
    PL_FMT_CAP_RENDERABLE = 1 << 3, PL_FMT_CAP_BLENDABLE = 1 << 4, PL_FMT_CAP_BLITTABLE = 1 << 5,
    PL_FMT_CAP_VERTEX = 1 << 6, PL_FMT_CAP_TEXEL_UNIFORM = 1 << 7,
    PL_FMT_CAP_TEXEL_STORAGE = 1 << 8, PL_FMT_CAP_HOST_READABLE = 1 << 9,
    PL_FMT_CAP_READWRITE = 1 << 10,
};
struct pl_fmt_t { /* gpu.h:270-340 */
    const char *name;
    uint64_t signature;
    enum pl_fmt_type type;
    enum pl_fmt_caps caps;
    int num_components;
    int component_depth[4];
    size_t internal_size;
    struct pl_fmt_plane { pl_fmt format; uint8_t shift_x, shift_y; } planes[4]; /* planar formats: none here */
    int num_planes;
    bool opaque;
    bool emulated;
    size_t texel_size;
    size_t texel_align;
    int host_bits[4];
    int sample_order[4];
    bool gatherable;
    const char *glsl_type;      /* GLSL / DRM interop members of the reference: NULL / 0 here */
    const char *glsl_format;
    uint32_t fourcc;
    const uint64_t *modifiers;
    int num_modifiers;
};
// gpu.h:325-331 — components stored in RGBA order / floating-point texels
PL_API bool pl_fmt_is_ordered(pl_fmt fmt);
PL_API bool pl_fmt_is_float(pl_fmt fmt);
PL_API pl_fmt pl_find_fmt(pl_gpu gpu, enum pl_fmt_type type, int num_components,
                          int min_depth, int host_bits, enum pl_fmt_caps caps);
PL_API pl_fmt pl_find_named_fmt(pl_gpu gpu, const char *name);

typedef struct pl_timer_t *pl_timer;
PL_API pl_timer pl_timer_create(pl_gpu gpu);
PL_API void pl_timer_destroy(pl_gpu gpu, pl_timer *);
PL_API uint64_t pl_timer_query(pl_gpu gpu, pl_timer);

/* External memory handles (gpu.h:158-217). Members exist so that the parameter structs have the
 * reference's layout; this backend imports CUDA device memory through pl_cuda_wrap() instead,
 * and pl_buf_create / pl_tex_create fail loudly when an export / import handle type is set. */
enum pl_handle_type {
    PL_HANDLE_FD = 1 << 0, PL_HANDLE_WIN32 = 1 << 1, PL_HANDLE_WIN32_KMT = 1 << 2,
    PL_HANDLE_DMA_BUF = 1 << 3, PL_HANDLE_HOST_PTR = 1 << 4, PL_HANDLE_MTL_TEX = 1 << 5,
    PL_HANDLE_IOSURFACE = 1 << 6,
};
union pl_handle { int fd; void *handle; void *ptr; };
struct pl_shared_mem {
    union pl_handle handle;
    size_t size, offset;
    uint64_t drm_format_mod;
    size_t stride_w, stride_h;
    unsigned plane;
};
enum pl_buf_mem_type { PL_BUF_MEM_AUTO = 0, PL_BUF_MEM_HOST, PL_BUF_MEM_DEVICE, PL_BUF_MEM_TYPE_COUNT };
struct pl_buf_params { /* gpu.h:441-520 */
    size_t size;
    bool host_writable, host_readable, host_mapped;
    bool uniform, storable, drawable;
    enum pl_buf_mem_type memory_type;
    pl_fmt format;
    enum pl_handle_type export_handle, import_handle;   /* must be 0 */
    struct pl_shared_mem shared_mem;
    const void *initial_data;
    void *user_data;
    pl_debug_tag debug_tag;
};
#define pl_buf_params(...) (&(struct pl_buf_params) { __VA_ARGS__ })
typedef const struct pl_buf_t { /* gpu.h:524-534 */
    struct pl_buf_params params;
    uint8_t *data;
    struct pl_shared_mem shared_mem;
} *pl_buf;
PL_API pl_buf pl_buf_create(pl_gpu gpu, const struct pl_buf_params *params);
PL_API void pl_buf_destroy(pl_gpu gpu, pl_buf *buf);
PL_API bool pl_buf_recreate(pl_gpu gpu, pl_buf *buf, const struct pl_buf_params *params);
PL_API void pl_buf_write(pl_gpu gpu, pl_buf buf, size_t buf_offset, const void *data, size_t size);
PL_API bool pl_buf_read(pl_gpu gpu, pl_buf buf, size_t buf_offset, void *dest, size_t size);
PL_API void pl_buf_copy(pl_gpu gpu, pl_buf dst, size_t dst_offset, pl_buf src, size_t src_offset, size_t size);
PL_API bool pl_buf_poll(pl_gpu gpu, pl_buf buf, uint64_t timeout);
/* src/gpu.h:145-165 — endian swapping between storable buffers (wordsize 2 or 4; src == dst
 * allowed for in-place). Library-internal in the reference, exported here because pl_upload_plane
 * uses it for `pl_plane_data.swapped` and the reference's conformance test calls it. */
struct pl_buf_copy_swap_params {
    pl_buf src;
    size_t src_offset;
    pl_buf dst;
    size_t dst_offset;
    size_t size;
    int wordsize;
};
PL_API bool pl_buf_copy_swap(pl_gpu gpu, const struct pl_buf_copy_swap_params *params);

enum pl_tex_sample_mode { PL_TEX_SAMPLE_NEAREST, PL_TEX_SAMPLE_LINEAR, PL_TEX_SAMPLE_MODE_COUNT };
enum pl_tex_address_mode { PL_TEX_ADDRESS_CLAMP, PL_TEX_ADDRESS_REPEAT, PL_TEX_ADDRESS_MIRROR, PL_TEX_ADDRESS_MODE_COUNT };
struct pl_tex_params { /* gpu.h:546-620 */
    int w, h, d;
    pl_fmt format;
    bool sampleable, renderable, storable, blit_src, blit_dst, host_writable, host_readable;
    enum pl_handle_type export_handle, import_handle;   /* must be 0 */
    struct pl_shared_mem shared_mem;
    const void *initial_data;
    void *user_data;
    pl_debug_tag debug_tag;
};
#define pl_tex_params(...) (&(struct pl_tex_params) { __VA_ARGS__ })
static inline int pl_tex_params_dimension(const struct pl_tex_params params)
{
    return params.d ? 3 : params.h ? 2 : 1;
}
enum pl_sampler_type { PL_SAMPLER_NORMAL, PL_SAMPLER_RECT, PL_SAMPLER_EXTERNAL, PL_SAMPLER_TYPE_COUNT };
typedef const struct pl_tex_t *pl_tex;
struct pl_tex_t { /* gpu.h:755-781 */
    struct pl_tex_params params;
    pl_tex planes[4];               /* planar formats: none on this backend */
    pl_tex parent;
    struct pl_shared_mem shared_mem;
    enum pl_sampler_type sampler_type;
};
PL_API pl_tex pl_tex_create(pl_gpu gpu, const struct pl_tex_params *params);
PL_API void pl_tex_destroy(pl_gpu gpu, pl_tex *tex);
PL_API bool pl_tex_recreate(pl_gpu gpu, pl_tex *tex, const struct pl_tex_params *params);
PL_API void pl_tex_invalidate(pl_gpu gpu, pl_tex tex);
PL_API void pl_tex_clear(pl_gpu gpu, pl_tex dst, const float color[4]);
// gpu.h:815-826 — clear with a typed colour; on this backend every clearable format is unorm /
// float (color.f); integer formats are rejected with an error message
union pl_clear_color {
    float f[4];
    int32_t i[4];
    uint32_t u[4];
};
PL_API void pl_tex_clear_ex(pl_gpu gpu, pl_tex dst, const union pl_clear_color color);
struct pl_tex_blit_params { /* gpu.h:820-845 */
    pl_tex src;             // needs params.blit_src
    pl_tex dst;             // needs params.blit_dst and a format of the same internal size
    pl_rect3d src_rc;       // region to copy (optional: whole texture); may be flipped
    pl_rect3d dst_rc;       // region to write (optional); the rest of dst is preserved
    enum pl_tex_sample_mode sample_mode; // used when the regions differ in size
};
#define pl_tex_blit_params(...) (&(struct pl_tex_blit_params) { __VA_ARGS__ })
PL_API void pl_tex_blit(pl_gpu gpu, const struct pl_tex_blit_params *params); /* gpu.h:849; 2D float / unorm textures */
struct pl_tex_transfer_params { /* gpu.h:852-919 */
    pl_tex tex;
    pl_rect3d rc;
    size_t row_pitch;
    size_t depth_pitch;
    pl_timer timer;
    void (*callback)(void *priv);
    void *priv;
    pl_buf buf;
    size_t buf_offset;
    void *ptr;
    bool no_import;
};
#define pl_tex_transfer_params(...) (&(struct pl_tex_transfer_params) { __VA_ARGS__ })
PL_API bool pl_tex_upload(pl_gpu gpu, const struct pl_tex_transfer_params *params);
PL_API bool pl_tex_download(pl_gpu gpu, const struct pl_tex_transfer_params *params);
PL_API bool pl_tex_poll(pl_gpu gpu, pl_tex tex, uint64_t timeout);
PL_API void pl_gpu_flush(pl_gpu gpu);
PL_API void pl_gpu_finish(pl_gpu gpu);
PL_API bool pl_gpu_is_failed(pl_gpu gpu);

enum pl_blend_mode { PL_BLEND_ZERO, PL_BLEND_ONE, PL_BLEND_SRC_ALPHA, PL_BLEND_ONE_MINUS_SRC_ALPHA, PL_BLEND_MODE_COUNT };
struct pl_blend_params { enum pl_blend_mode src_rgb, dst_rgb, src_alpha, dst_alpha; };

/* ---- the CUDA backend constructor (sibling of dummy.h:36-100 / vulkan.h) ------------------ */
struct pl_cuda_params {
    int device;        // CUDA device ordinal
    void *stream;      // cudaStream_t to order all work on (NULL = create a private one)
    bool no_timers;    // skip cudaEvent timers in pl_dispatch
    bool no_async_transfers; // run pl_tex_upload/download on the main stream instead of the
                             // copy-engine streams (default: overlap transfers with kernels)
    // Sampling positions: by default they are evaluated exactly (like a rasteriser's
    // interpolators). Set this to reproduce, in fp32, the attribute emulation the reference
    // uses for compute shaders (src/dispatch.c:1038-1062); it disables the tabulated-phase
    // fast paths of the scalers. See DESIGN.md "coordinates".
    bool literal_coords;
};
#define pl_cuda_params(...) (&(struct pl_cuda_params) { __VA_ARGS__ })
PL_API pl_gpu pl_cuda_create(pl_log log, const struct pl_cuda_params *params);
PL_API void pl_cuda_destroy(pl_gpu *gpu);
PL_API void *pl_cuda_stream(pl_gpu gpu); // the cudaStream_t all work is ordered on
// Wrap device memory that the caller owns (e.g. an NVDEC surface or a framework tensor) as a
// pl_tex without copying; counterpart of pl_vulkan_wrap (vulkan.h). pitch in bytes.
// Alignment: `ptr` and `pitch` must be multiples of 16 bytes (the kernels access texels as whole
// vectors and store runs of four texels as 16-byte words); anything from cudaMalloc /
// cudaMallocPitch / NVDEC qualifies. Misaligned surfaces are rejected (NULL).
struct pl_cuda_wrap_params {
    void *ptr;
    size_t pitch;
    int w, h;
    pl_fmt format;
};
#define pl_cuda_wrap_params(...) (&(struct pl_cuda_wrap_params) { __VA_ARGS__ })
PL_API pl_tex pl_cuda_wrap(pl_gpu gpu, const struct pl_cuda_wrap_params *params);
PL_API void *pl_cuda_tex_ptr(pl_tex tex, size_t *out_pitch);
// Number of kernels launched through this gpu since creation (bench.py: gpu_launches)
PL_API uint64_t pl_cuda_launch_count(pl_gpu gpu);
// Makes the main stream (pl_cuda_stream) wait for every transfer enqueued so far on the
// copy-engine streams — e.g. before recording an event that should cover a whole frame.
PL_API void pl_cuda_join(pl_gpu gpu);
/* Copy up to 256 KiB between device memory and/or pinned host memory with a kernel on `stream`
 * (a cudaStream_t; NULL = pl_cuda_stream(gpu)) instead of a DMA engine: small control payloads
 * (peak statistics, LUTs) do not queue behind frame-sized transfers of the copy engines. The
 * host side must be pinned (cudaMallocHost / cudaHostRegister); completion = stream order. */
PL_API bool pl_cuda_small_copy(pl_gpu gpu, void *dst, const void *src, size_t size, void *stream);
// Diagnostics: evaluates the kernels' deterministic transcendental functions (plb_math.cuh) on
// host arrays of n floats. fn: 0 pow(a,b) 1 exp(a) 2 log(a) 3 sin(a) 4 cos(a) 5 atan2(a,b)
// (binary64 definitions); 6 sin 7 cos 8 atan2 (fp32 hot-path definitions); 16 + id: transfer curve
// `id` of pl_cuda_curve_table with its exact fallback outside the tabulated domain.
// Returns false on a CUDA error. Lets integrators check cross-machine reproducibility.
PL_API bool pl_cuda_math_eval(pl_gpu gpu, int fn, const float *a, const float *b, float *out,
                              size_t n);
// Diagnostics (host only, no GPU needed): the piecewise-cubic table that defines transfer curve
// `id` (csrc/curves.h: PQ EOTF / OETF, BT.1886 powers) exactly as the kernels receive it.
// Segment 0 is the constant below the domain; segment i >= 1 covers the float bit patterns
// [(base + i) << (23 - mbits), (base + i + 1) << (23 - mbits)). `coef` may be NULL to query the
// sizes, else it receives 4 * (*nseg + 1) floats (c0..c3 per segment).
PL_API bool pl_cuda_curve_table(int id, int *base, int *nseg, int *mbits, float *coef);

/* ---- shaders.h -------------------------------------------------------------------------- */
typedef struct pl_shader_t *pl_shader;
struct pl_shader_params { /* shaders.h:36-70 */
    uint8_t id;
    pl_gpu gpu;
    uint8_t index;
    struct pl_glsl_version glsl;
    bool dynamic_constants;
};
#define pl_shader_params(...) (&(struct pl_shader_params) { __VA_ARGS__ })
PL_API pl_shader pl_shader_alloc(pl_log log, const struct pl_shader_params *params);
PL_API void pl_shader_free(pl_shader *sh);
PL_API void pl_shader_reset(pl_shader sh, const struct pl_shader_params *params);
PL_API bool pl_shader_is_failed(const pl_shader sh);
PL_API bool pl_shader_is_compute(const pl_shader sh);
PL_API bool pl_shader_output_size(const pl_shader sh, int *w, int *h);
typedef struct pl_shader_obj_t *pl_shader_obj;
PL_API void pl_shader_obj_destroy(pl_shader_obj *obj);

/* ---- shaders/sampling.h ----------------------------------------------------------------- */
struct pl_sample_src { /* sampling.h:34-62 */
    pl_tex tex;
    pl_rect2df rect;
    enum pl_tex_address_mode address_mode;
    int tex_w, tex_h;
    enum pl_fmt_type format;
    enum pl_sampler_type sampler;
    enum pl_tex_sample_mode mode;
    float sampled_w, sampled_h;
    int components;
    uint8_t component_mask;
    int new_w, new_h;
    float scale;
};
#define pl_sample_src(...) (&(struct pl_sample_src) { __VA_ARGS__ })
struct pl_deband_params { /* sampling.h:66-103 */
    int iterations;
    float threshold;
    float radius;
    float grain;
    float grain_neutral[3];
};
#define PL_DEBAND_DEFAULTS .iterations = 1, .threshold = 3.0, .radius = 16.0, .grain = 4.0,
#define pl_deband_params(...) (&(struct pl_deband_params) {PL_DEBAND_DEFAULTS __VA_ARGS__ })
PL_API extern const struct pl_deband_params pl_deband_default_params;
PL_API void pl_shader_deband(pl_shader sh, const struct pl_sample_src *src,
                             const struct pl_deband_params *params);
PL_API bool pl_shader_sample_direct(pl_shader sh, const struct pl_sample_src *src);
PL_API bool pl_shader_sample_nearest(pl_shader sh, const struct pl_sample_src *src);
PL_API bool pl_shader_sample_bilinear(pl_shader sh, const struct pl_sample_src *src);
/* built-in fast scalers (sampling.h:120-150): a few bilinear fetches, for up-scaling */
PL_API bool pl_shader_sample_bicubic(pl_shader sh, const struct pl_sample_src *src);
PL_API bool pl_shader_sample_hermite(pl_shader sh, const struct pl_sample_src *src);
PL_API bool pl_shader_sample_gaussian(pl_shader sh, const struct pl_sample_src *src);
PL_API bool pl_shader_sample_oversample(pl_shader sh, const struct pl_sample_src *src, float threshold);
struct pl_sample_filter_params { /* sampling.h:153-178 */
    struct pl_filter_config filter;
    float antiring;
    bool no_compute;
    bool no_widening;
    pl_shader_obj *lut;
    int lut_entries;        /* deprecated in the reference (fixed at 256); kept for layout */
    float cutoff;           /* deprecated in the reference (fixed at 1e-3); kept for layout */
};
#define pl_sample_filter_params(...) (&(struct pl_sample_filter_params) { __VA_ARGS__ })
PL_API bool pl_shader_sample_polar(pl_shader sh, const struct pl_sample_src *src,
                                   const struct pl_sample_filter_params *params);
PL_API bool pl_shader_sample_ortho2(pl_shader sh, const struct pl_sample_src *src,
                                    const struct pl_sample_filter_params *params);

/* ---- shaders/colorspace.h --------------------------------------------------------------- */
PL_API void pl_shader_set_alpha(pl_shader sh, struct pl_color_repr *repr, enum pl_alpha_mode mode);
PL_API void pl_shader_dovi_reshape(pl_shader sh, const struct pl_dovi_metadata *data);
PL_API void pl_shader_decode_color(pl_shader sh, struct pl_color_repr *repr,
                                   const struct pl_color_adjustment *params);
PL_API void pl_shader_encode_color(pl_shader sh, const struct pl_color_repr *repr);
PL_API void pl_shader_linearize(pl_shader sh, const struct pl_color_space *csp);
/* shaders/colorspace.h:185-189: perceptual intensity of the (non-linear) input colour in `csp` */
PL_API void pl_shader_extract_features(pl_shader sh, struct pl_color_space csp);
PL_API void pl_shader_delinearize(pl_shader sh, const struct pl_color_space *csp);
/* shaders/lut.h:30-78 — custom 1-D / 3-D LUTs (.cube files). `data` holds size[0] (x size[1] x
 * size[2]) RGB triples, red fastest; all-zero shaper matrices are skipped. */
struct pl_custom_lut {
    uint64_t signature;
    int size[3];
    const float *data;
    pl_matrix3x3 shaper_in, shaper_out;
    struct pl_color_repr repr_in, repr_out;
    struct pl_color_space color_in, color_out;
};
PL_API struct pl_custom_lut *pl_lut_parse_cube(pl_log log, const char *str, size_t str_len);
PL_API void pl_lut_free(struct pl_custom_lut **lut);
PL_API void pl_shader_custom_lut(pl_shader sh, const struct pl_custom_lut *lut,
                                 pl_shader_obj *lut_state);
struct pl_sigmoid_params { float center, slope; };
#define PL_SIGMOID_DEFAULTS .center = 0.75, .slope = 6.50,
PL_API extern const struct pl_sigmoid_params pl_sigmoid_default_params;
PL_API void pl_shader_sigmoidize(pl_shader sh, const struct pl_sigmoid_params *params);
PL_API void pl_shader_unsigmoidize(pl_shader sh, const struct pl_sigmoid_params *params);
struct pl_peak_detect_params { /* shaders/colorspace.h:104-151 */
    float smoothing_period;
    float scene_threshold_low;
    float scene_threshold_high;
    float percentile;
    float black_cutoff;
    bool allow_delayed;
    float minimum_peak;     /* deprecated in the reference; kept for layout */
};
#define PL_PEAK_DETECT_DEFAULTS \
    .smoothing_period = 20.0f, .scene_threshold_low = 1.0f, .scene_threshold_high = 3.0f, \
    .percentile = 100.0f, .black_cutoff = 1.0f,
#define PL_PEAK_DETECT_HQ_DEFAULTS PL_PEAK_DETECT_DEFAULTS .percentile = 99.995f,
#define pl_peak_detect_params(...) (&(struct pl_peak_detect_params) { PL_PEAK_DETECT_DEFAULTS __VA_ARGS__ })
PL_API extern const struct pl_peak_detect_params pl_peak_detect_default_params;
PL_API extern const struct pl_peak_detect_params pl_peak_detect_high_quality_params;
PL_API bool pl_shader_detect_peak(pl_shader sh, struct pl_color_space csp, pl_shader_obj *state,
                                  const struct pl_peak_detect_params *params);
PL_API bool pl_get_detected_hdr_metadata(const pl_shader_obj state, struct pl_hdr_metadata *metadata);
PL_API void pl_reset_detected_peak(pl_shader_obj state);
struct pl_color_map_params { /* shaders/colorspace.h:227-312 */
    const struct pl_gamut_map_function *gamut_mapping;
    struct pl_gamut_map_constants gamut_constants;
    int lut3d_size[3];
    bool lut3d_tricubic;
    bool gamut_expansion;
    const struct pl_tone_map_function *tone_mapping_function;
    struct pl_tone_map_constants tone_constants;
    bool inverse_tone_mapping;
    enum pl_hdr_metadata_type metadata;
    int lut_size;
    float contrast_recovery;
    float contrast_smoothness;
    bool force_tone_mapping_lut;
    bool visualize_lut;
    pl_rect2df visualize_rect;
    float visualize_hue;
    float visualize_theta;
    bool show_clipping;
    /* members the reference deprecated (shaders/colorspace.h:306-311); ignored, kept for layout */
    int tone_mapping_mode;
    float tone_mapping_param;
    float tone_mapping_crosstalk;
    int intent;
    int gamut_mode;
    float hybrid_mix;
};
#define PL_COLOR_MAP_DEFAULTS \
    .gamut_mapping = &pl_gamut_map_perceptual, .tone_mapping_function = &pl_tone_map_spline, \
    .gamut_constants = { PL_GAMUT_MAP_CONSTANTS }, .tone_constants = { PL_TONE_MAP_CONSTANTS }, \
    .metadata = PL_HDR_METADATA_ANY, .lut3d_size = {48, 32, 256}, .lut_size = 256, \
    .visualize_rect = {0, 0, 1, 1}, .contrast_smoothness = 3.5f,
#define PL_COLOR_MAP_HQ_DEFAULTS PL_COLOR_MAP_DEFAULTS .contrast_recovery = 0.30f,
#define pl_color_map_params(...) (&(struct pl_color_map_params) { PL_COLOR_MAP_DEFAULTS __VA_ARGS__ })
PL_API extern const struct pl_color_map_params pl_color_map_default_params;
PL_API extern const struct pl_color_map_params pl_color_map_high_quality_params;
struct pl_color_map_args {
    struct pl_color_space src;
    struct pl_color_space dst;
    bool prelinearized;
    pl_shader_obj *state;
    pl_tex feature_map;
};
#define pl_color_map_args(...) (&(struct pl_color_map_args) { __VA_ARGS__ })
PL_API void pl_shader_color_map_ex(pl_shader sh, const struct pl_color_map_params *params,
                                   const struct pl_color_map_args *args);
PL_API void pl_shader_color_map(pl_shader sh, const struct pl_color_map_params *params,
                                struct pl_color_space src, struct pl_color_space dst,
                                pl_shader_obj *state, bool prelinearized);

/* ---- shaders/dithering.h ---------------------------------------------------------------- */
enum pl_dither_method {
    PL_DITHER_BLUE_NOISE, PL_DITHER_ORDERED_LUT, PL_DITHER_ORDERED_FIXED, PL_DITHER_WHITE_NOISE,
    PL_DITHER_METHOD_COUNT,
};
struct pl_dither_params {
    enum pl_dither_method method;
    int lut_size;
    bool temporal;
    enum pl_color_transfer transfer;
};
#define PL_DITHER_DEFAULTS .method = PL_DITHER_BLUE_NOISE, .lut_size = 6, .temporal = false,
#define pl_dither_params(...) (&(struct pl_dither_params) { PL_DITHER_DEFAULTS __VA_ARGS__ })
PL_API extern const struct pl_dither_params pl_dither_default_params;
PL_API void pl_shader_dither(pl_shader sh, int new_depth, pl_shader_obj *dither_state,
                             const struct pl_dither_params *params);
struct pl_error_diffusion_params {
    pl_tex input_tex;
    pl_tex output_tex;
    int new_depth;
    const struct pl_error_diffusion_kernel *kernel;
};
#define pl_error_diffusion_params(...) (&(struct pl_error_diffusion_params) { __VA_ARGS__ })
PL_API size_t pl_error_diffusion_shmem_req(const struct pl_error_diffusion_kernel *kernel, int height);
PL_API bool pl_shader_error_diffusion(pl_shader sh, const struct pl_error_diffusion_params *params);

/* ---- shaders/film_grain.h --------------------------------------------------------------- */
enum pl_film_grain_type { PL_FILM_GRAIN_NONE = 0, PL_FILM_GRAIN_AV1, PL_FILM_GRAIN_H274, PL_FILM_GRAIN_COUNT };
struct pl_av1_grain_data { /* film_grain.h:45-66 */
    int num_points_y;
    uint8_t points_y[14][2];
    bool chroma_scaling_from_luma;
    int num_points_uv[2];
    uint8_t points_uv[2][10][2];
    int scaling_shift;
    int ar_coeff_lag;
    int8_t ar_coeffs_y[24];
    int8_t ar_coeffs_uv[2][25];
    int ar_coeff_shift;
    int grain_scale_shift;
    int8_t uv_mult[2];
    int8_t uv_mult_luma[2];
    int16_t uv_offset[2];
    bool overlap;
};
struct pl_h274_grain_data { /* film_grain.h:70-86 */
    int model_id;
    int blending_mode_id;
    int log2_scale_factor;
    bool component_model_present[3];
    uint16_t num_intensity_intervals[3];
    uint8_t num_model_values[3];
    const uint8_t *intensity_interval_lower_bound[3];
    const uint8_t *intensity_interval_upper_bound[3];
    const int16_t (*comp_model_value[3])[6];
};
struct pl_film_grain_data {
    enum pl_film_grain_type type;
    uint64_t seed;
    union {
        struct pl_av1_grain_data av1;
        struct pl_h274_grain_data h274;
    } params;
};
struct pl_film_grain_params { /* film_grain.h:100-128 */
    struct pl_film_grain_data data;
    pl_tex tex;
    struct pl_color_repr *repr;
    int components;
    int component_mapping[4];
    pl_tex luma_tex;
    int luma_comp;
};
#define pl_film_grain_params(...) (&(struct pl_film_grain_params) { __VA_ARGS__ })
PL_API bool pl_needs_film_grain(const struct pl_film_grain_params *params);
// Diagnostics (host only, no GPU needed): the AV1 tables pl_shader_film_grain generates for
// `data` / `repr` — grain_y[64*64], grain_cb and grain_cr [(64 >> sub_x) * (64 >> sub_y)],
// offsets[(align(tex_w << sub_x, 128) / 32) * (align(tex_h << sub_y, 128) / 32)], scaling[3][256].
// Any output pointer may be NULL.
PL_API bool pl_cuda_av1_tables_host(const struct pl_film_grain_data *data,
                                    const struct pl_color_repr *repr, int sub_x, int sub_y,
                                    int tex_w, int tex_h, float *grain_y, float *grain_cb,
                                    float *grain_cr, uint32_t *offsets, float *scaling);
// Diagnostics (host only): the H.274 grain data base, out[(13*64) * (13*64)] floats.
PL_API bool pl_cuda_h274_database_host(float *out);
PL_API bool pl_shader_film_grain(pl_shader sh, pl_shader_obj *grain_state,
                                 const struct pl_film_grain_params *params);

/* ---- dispatch.h ------------------------------------------------------------------------- */
typedef struct pl_dispatch_t *pl_dispatch;
PL_API pl_dispatch pl_dispatch_create(pl_log log, pl_gpu gpu);
PL_API void pl_dispatch_destroy(pl_dispatch *dp);
PL_API void pl_dispatch_reset_frame(pl_dispatch dp);
PL_API pl_shader pl_dispatch_begin(pl_dispatch dp);
/* shaders.h:121-138; refcounting is a no-op here: the info lives as long as the pl_dispatch */
typedef const struct pl_shader_info_t {
    struct pl_shader_params params;
    const char **steps;
    int num_steps;
    const char *description;
} *pl_shader_info;
PL_API pl_shader_info pl_shader_info_ref(pl_shader_info info);
PL_API void pl_shader_info_deref(pl_shader_info *info);

struct pl_dispatch_info { /* dispatch.h:50-65 */
    pl_shader_info shader;
    uint64_t signature;
    uint64_t samples[256];
    int num_samples;
    uint64_t last;
    uint64_t peak;
    uint64_t average;
};
/* dispatch.h:77-82: `cb` runs for every successfully dispatched shader; the execution times it
 * reports are cudaEvent intervals on the gpu's stream and arrive a few dispatches late (when
 * the events have completed), never by blocking the stream. */
PL_API void pl_dispatch_callback(pl_dispatch dp, void *priv,
                                 void (*cb)(void *priv, const struct pl_dispatch_info *));

struct pl_dispatch_params { /* dispatch.h:84-115 */
    pl_shader *shader;
    pl_tex target;
    pl_rect2d rect;
    const struct pl_blend_params *blend_params;
    pl_timer timer;
};
#define pl_dispatch_params(...) (&(struct pl_dispatch_params) { __VA_ARGS__ })
PL_API bool pl_dispatch_finish(pl_dispatch dp, const struct pl_dispatch_params *params);
struct pl_dispatch_compute_params {
    pl_shader *shader;
    int dispatch_size[3];
    int width, height;
    pl_timer timer;
};
#define pl_dispatch_compute_params(...) (&(struct pl_dispatch_compute_params) { __VA_ARGS__ })
PL_API bool pl_dispatch_compute(pl_dispatch dp, const struct pl_dispatch_compute_params *params);
PL_API void pl_dispatch_abort(pl_dispatch dp, pl_shader *sh);
// Overrides the frame counter that seeds temporal PRNGs/dither (dispatch.c:1613-1622); a
// frame-parallel multi-GPU run feeds the GLOBAL frame number here.
PL_API void pl_dispatch_set_frame_index(pl_dispatch dp, uint8_t index);

/* ---- renderer.h ------------------------------------------------------------------------- */
typedef struct pl_renderer_t *pl_renderer;
enum pl_render_error { /* renderer.h:41-56 */
    PL_RENDER_ERR_NONE = 0, PL_RENDER_ERR_FBO = 1 << 0, PL_RENDER_ERR_SAMPLING = 1 << 1,
    PL_RENDER_ERR_DEBANDING = 1 << 2, PL_RENDER_ERR_BLENDING = 1 << 3,
    PL_RENDER_ERR_OVERLAY = 1 << 4, PL_RENDER_ERR_PEAK_DETECT = 1 << 5,
    PL_RENDER_ERR_FILM_GRAIN = 1 << 6, PL_RENDER_ERR_FRAME_MIXING = 1 << 7,
    PL_RENDER_ERR_DEINTERLACING = 1 << 8, PL_RENDER_ERR_ERROR_DIFFUSION = 1 << 9,
    PL_RENDER_ERR_HOOKS = 1 << 10, PL_RENDER_ERR_CONTRAST_RECOVERY = 1 << 11,
    PL_RENDER_ERR_BLUR = 1 << 12,
};
struct pl_render_errors {
    enum pl_render_error errors;
    const uint64_t *disabled_hooks;
    int num_disabled_hooks;
};
PL_API pl_renderer pl_renderer_create(pl_log log, pl_gpu gpu);
PL_API void pl_renderer_destroy(pl_renderer *rr);
PL_API struct pl_render_errors pl_renderer_get_errors(pl_renderer rr);
PL_API void pl_renderer_reset_errors(pl_renderer rr, const struct pl_render_errors *errors);
enum pl_lut_type { PL_LUT_UNKNOWN = 0, PL_LUT_NATIVE, PL_LUT_NORMALIZED, PL_LUT_CONVERSION };
enum pl_clear_mode { PL_CLEAR_COLOR = 0, PL_CLEAR_TILES, PL_CLEAR_SKIP, PL_CLEAR_BLUR, PL_CLEAR_MODE_COUNT };
enum pl_render_stage { /* renderer.h:106-113 */
    PL_RENDER_STAGE_FRAME,
    PL_RENDER_STAGE_BLEND,
    PL_RENDER_STAGE_COUNT,
};
struct pl_render_info { /* renderer.h:115-126 */
    const struct pl_dispatch_info *pass;
    enum pl_render_stage stage;
    int index;
    int count;
};
struct pl_render_params { /* renderer.h:130-368; members outside the hot path are accepted
                             but must be NULL/0, otherwise pl_render_image fails loudly */
    const struct pl_filter_config *upscaler;
    const struct pl_filter_config *downscaler;
    const struct pl_filter_config *plane_upscaler;
    const struct pl_filter_config *plane_downscaler;
    float antiringing_strength;
    const struct pl_filter_config *frame_mixer;
    const struct pl_deband_params *deband_params;
    const struct pl_sigmoid_params *sigmoid_params;
    const struct pl_color_adjustment *color_adjustment;
    const struct pl_peak_detect_params *peak_detect_params;
    const struct pl_color_map_params *color_map_params;
    const struct pl_dither_params *dither_params;
    const struct pl_error_diffusion_kernel *error_diffusion;
    const void *cone_params;
    const struct pl_blend_params *blend_params;
    const struct pl_deinterlace_params *deinterlace_params;   /* renderer.h:305-310 */
    const void *distort_params;
    const void * const *hooks;
    int num_hooks;
    const struct pl_custom_lut *lut;    /* renderer.h:259-270 */
    enum pl_lut_type lut_type;
    enum pl_clear_mode background;
    enum pl_clear_mode border;
    float background_color[3];
    float background_transparency;
    float tile_colors[2][3];
    int tile_size;
    float blur_radius;
    float corner_rounding;
    bool skip_anti_aliasing;
    bool preserve_mixing_cache;
    bool skip_caching_single_frame;
    bool disable_linear_scaling;
    bool disable_builtin_scalers;
    bool correct_subpixel_offsets;
    bool force_dither;
    bool disable_dither_gamma_correction;
    bool disable_fbos;
    bool force_low_bit_depth_fbos;
    bool dynamic_constants;
    void (*info_callback)(void *priv, const struct pl_render_info *info);
    void *info_priv;
    /* members the reference deprecated (renderer.h:361-367); ignored, kept for layout */
    bool allow_delayed_peak_detect;
    const void *icc_params;
    bool ignore_icc_profiles;
    int lut_entries;
    float polar_cutoff;
    bool skip_target_clearing;
    bool blend_against_tiles;
};
#define PL_RENDER_DEFAULTS \
    .color_map_params = &pl_color_map_default_params, \
    .color_adjustment = &pl_color_adjustment_neutral, \
    .tile_colors = {{0.93, 0.93, 0.93}, {0.87, 0.87, 0.87}}, .tile_size = 32, .blur_radius = 16.0,
#define pl_render_params(...) (&(struct pl_render_params) { PL_RENDER_DEFAULTS __VA_ARGS__ })
PL_API extern const struct pl_render_params pl_render_fast_params;
PL_API extern const struct pl_render_params pl_render_default_params;
PL_API extern const struct pl_render_params pl_render_high_quality_params;
#define PL_MAX_PLANES 4
struct pl_plane { /* renderer.h:404-472 */
    pl_tex texture;
    enum pl_tex_address_mode address_mode;
    bool flipped;
    int components;
    int component_mapping[4];
    float shift_x, shift_y;
};
enum pl_field { PL_FIELD_NONE = 0, PL_FIELD_EVEN, PL_FIELD_ODD,
                PL_FIELD_TOP = PL_FIELD_EVEN, PL_FIELD_BOTTOM = PL_FIELD_ODD };
/* shaders/deinterlacing.h:30-137 — deinterlacing of interleaved (MBAFF) frames */
struct pl_field_pair { pl_tex top; };
#define pl_field_pair(...) ((struct pl_field_pair) { __VA_ARGS__ })
struct pl_deinterlace_source {
    struct pl_field_pair prev, cur, next;   /* prev / next: optional neighbouring frames */
    enum pl_field field;                    /* the field being shown (its lines are kept) */
    enum pl_field first_field;              /* temporal order of the fields; NONE = top first */
    uint8_t component_mask;                 /* 0 = all components of the texture */
};
#define pl_deinterlace_source(...) (&(struct pl_deinterlace_source) { __VA_ARGS__ })
enum pl_deinterlace_algorithm {
    PL_DEINTERLACE_WEAVE = 0, PL_DEINTERLACE_BOB, PL_DEINTERLACE_YADIF, PL_DEINTERLACE_BWDIF,
    PL_DEINTERLACE_ALGORITHM_COUNT,
};
static inline bool pl_deinterlace_needs_refs(enum pl_deinterlace_algorithm algo)
{
    return algo >= PL_DEINTERLACE_YADIF;
}
struct pl_deinterlace_params {
    enum pl_deinterlace_algorithm algo;
    bool skip_spatial_check;
};
#define PL_DEINTERLACE_DEFAULTS .algo = PL_DEINTERLACE_BWDIF,
#define pl_deinterlace_params(...) (&(struct pl_deinterlace_params) { PL_DEINTERLACE_DEFAULTS __VA_ARGS__ })
PL_API extern const struct pl_deinterlace_params pl_deinterlace_default_params;
PL_API void pl_shader_deinterlace(pl_shader sh, const struct pl_deinterlace_source *src,
                                  const struct pl_deinterlace_params *params);
struct pl_frame { /* renderer.h:528-655 */
    int num_planes;
    struct pl_plane planes[PL_MAX_PLANES];
    enum pl_field field;
    enum pl_field first_field;
    const struct pl_frame *prev, *next;
    bool (*acquire)(pl_gpu gpu, struct pl_frame *frame);
    void (*release)(pl_gpu gpu, struct pl_frame *frame);
    struct pl_color_repr repr;
    struct pl_color_space color;
    const void *icc;            /* pl_icc_object: must be NULL (ICC is outside the hot path) */
    struct pl_icc_profile { const void *data; size_t len; uint64_t signature; } profile; /* must be empty */
    const struct pl_custom_lut *lut;    /* renderer.h:590-600 */
    enum pl_lut_type lut_type;
    pl_rect2df crop;
    pl_rotation rotation;
    float pixel_aspect_ratio;
    const void *overlays;
    int num_overlays;
    struct pl_film_grain_data film_grain;
    void *user_data;
};
PL_API void pl_frame_set_chroma_location(struct pl_frame *frame, enum pl_chroma_location chroma_loc);
PL_API bool pl_frame_is_cropped(const struct pl_frame *frame);
/* renderer.h:675-687: reset every plane of a frame to an sRGB colour (non-premultiplied), converted
 * to the frame's colour space and representation */
PL_API void pl_frame_clear_rgba(pl_gpu gpu, const struct pl_frame *frame, const float clear_color[4]);
static inline void pl_frame_clear(pl_gpu gpu, const struct pl_frame *frame, const float clear_color[3])
{
    const float rgba[4] = { clear_color[0], clear_color[1], clear_color[2], 1.0f };
    pl_frame_clear_rgba(gpu, frame, rgba);
}
PL_API void pl_frames_infer(pl_renderer rr, struct pl_frame *image, struct pl_frame *target);
PL_API bool pl_render_image(pl_renderer rr, const struct pl_frame *image,
                            const struct pl_frame *target, const struct pl_render_params *params);

/* renderer.h:754-847 — frame mixing: several source frames around one output vsync. `timestamps`
 * are relative to the target vsync, in units of source frames; `signatures` identify frames
 * across calls (their rendered versions are cached). pl_render_image_mix renders every frame
 * that contributes (weight from `pl_render_params.frame_mixer`: NULL = only the nearest frame,
 * pl_filter_function_oversample kernel = visible fraction of the vsync interval, any other filter
 * = pl_filter_sample at the timestamp) through the whole path into a cached rgba16hf image,
 * mixes them in linear light and sends the result through the output stage.
 * Differences on this backend: `preserve_mixing_cache` is treated as false. */
struct pl_frame_mix {
    int num_frames;
    const struct pl_frame **frames;
    const uint64_t *signatures;
    const float *timestamps;
    float vsync_duration;
};
PL_API const struct pl_frame *pl_frame_mix_current(const struct pl_frame_mix *mix);
PL_API const struct pl_frame *pl_frame_mix_nearest(const struct pl_frame_mix *mix);
PL_API void pl_frames_infer_mix(pl_renderer rr, const struct pl_frame_mix *mix,
                                struct pl_frame *target, struct pl_frame *out_ref);
PL_API bool pl_render_image_mix(pl_renderer rr, const struct pl_frame_mix *images,
                                const struct pl_frame *target, const struct pl_render_params *params);
PL_API void pl_renderer_flush_cache(pl_renderer rr);

/* ---- utils/upload.h: frame ingest (src/utils/upload.c) ------------------------------------ */
struct pl_plane_data { /* utils/upload.h:33-73 */
    enum pl_fmt_type type;      /* UNORM / SNORM / FLOAT */
    int width, height;
    int component_size[4];      /* bits per component, in memory order */
    int component_pad[4];       /* ignored bits in front of each component */
    int component_map[4];       /* colour channel (0..3) each component carries */
    size_t pixel_stride;        /* bytes between pixels */
    size_t row_stride;          /* bytes between rows (0 = tightly packed) */
    bool swapped;               /* samples are in the non-native byte order */
    const void *pixels;         /* either host memory ... */
    pl_buf buf;                 /* ... or a buffer + offset */
    size_t buf_offset;
    void (*callback)(void *priv);
    void *priv;
};
PL_API void pl_plane_data_from_mask(struct pl_plane_data *data, uint64_t mask[4]);
PL_API void pl_plane_data_from_comps(struct pl_plane_data *data, int size[4], int shift[4]);
PL_API bool pl_plane_data_align(struct pl_plane_data *data, struct pl_bit_encoding *out_bits);
PL_API pl_fmt pl_plane_find_fmt(pl_gpu gpu, int out_map[4], const struct pl_plane_data *data);
PL_API bool pl_upload_plane(pl_gpu gpu, struct pl_plane *out_plane, pl_tex *tex,
                            const struct pl_plane_data *data);
PL_API bool pl_recreate_plane(pl_gpu gpu, struct pl_plane *out_plane, pl_tex *tex,
                              const struct pl_plane_data *data);
/* Device-less object carrying only this backend's format table and limits, for format queries
 * (pl_find_fmt, pl_find_named_fmt, pl_plane_find_fmt) without a GPU. Nothing else works on it. */
PL_API pl_gpu pl_cuda_format_catalogue(void);
PL_API bool pl_renderer_get_hdr_metadata(pl_renderer rr, struct pl_hdr_metadata *metadata);

/* ---- frame-parallel multi-GPU support (no reference counterpart) ------------
 * The only cross-frame state of the path is the peak-detection IIR (shaders/colorspace.c:
 * 1124-1152). A rank that renders a strided subset of a stream measures its frames with
 * pl_renderer_measure_peak() (kernel only, raw statistics = struct peak_buf_data, 3264 B),
 * the ranks exchange the statistics (NCCL all-gather, done by the caller), every rank
 * replays the recurrence in frame order with pl_peak_state_update(), and renders frame N
 * with pl_renderer_set_peak_state() holding the smoothed values the reference would have had. */
#define PL_PEAK_BUF_SIZE 3264
struct pl_peak_state { float avg_pq, max_pq; };
PL_API bool pl_renderer_measure_peak(pl_renderer rr, const struct pl_frame *image,
                                     const struct pl_render_params *params,
                                     void *out_peak_buf /* PL_PEAK_BUF_SIZE bytes, host */);
/* Same, but the statistics stay on the device: `dev_peak_buf` (PL_PEAK_BUF_SIZE bytes of device
 * memory, e.g. the send buffer of an NCCL all-gather) is filled by an asynchronous copy on the
 * gpu's stream; the host does not wait. Up to four measured images may be pending (measured
 * ahead of their pl_render_image), which lets a pipeline overlap the exchange of frame N's
 * statistics with the measurement of frame N+1. */
PL_API bool pl_renderer_measure_peak_dev(pl_renderer rr, const struct pl_frame *image,
                                         const struct pl_render_params *params,
                                         void *dev_peak_buf /* PL_PEAK_BUF_SIZE bytes, device */);
PL_API bool pl_peak_state_update(struct pl_peak_state *state, const void *peak_buf,
                                 const struct pl_peak_detect_params *params);
PL_API void pl_renderer_set_peak_state(pl_renderer rr, const struct pl_peak_state *state);
PL_API void pl_renderer_set_frame_index(pl_renderer rr, uint8_t index);

PL_API_END
#endif /* LIBPLACEBO_B200_H_ */
