/*
 * TEST INFRASTRUCTURE — CPU oracle for the libplacebo per-frame hot path.
 *
 * This is a plain-C fp32 restatement of the GLSL bodies that the reference generates in
 * src/shaders/{sampling,colorspace,dithering,film_grain_av1,film_grain_h274}.c, plus the
 * conventions of src/dispatch.c (pixel coordinates) and src/shaders/lut.c (LUT addressing).
 * It is the checker for the CUDA kernels in libplacebo_b200/csrc and the "port" CPU baseline
 * of bench.py. It is NOT shipped and nothing under libplacebo_b200/ may include or link it.
 *
 * Pinning status: the reference's pixel arithmetic exists only as
 * GLSL that cannot execute in this environment, and its own tests do not pin pixel outputs of
 * polar/ortho/deband/grain/dither (src/tests/gpu_tests.c:912-997 are run-only). So:
 *   - every PARAMETER the oracle consumes (filter LUTs, tone/gamut LUTs, matrices, dither
 *     matrices, grain tables) comes from the real reference code via oracle/harvest.c;
 *   - functions with reference KATs are pinned against them in tests/test_oracle_kat.py
 *     (peak detection vs CPU mean/max: gpu_tests.c:753-785; TRC round trips :561-577;
 *     scRGB gamut goldens :589-667; colour-system round trips :673-703; filter tests
 *     tests/filters.c:20-72);
 *   - the remaining pixel bodies are "parity unpinned by reference tests": restated line by
 *     line from the harvested GLSL.
 *
 * Arithmetic conventions (GLSL leaves these implementation-defined; we fix them):
 *   - all pixel math in IEEE binary32, compiled with -ffp-contract=off; `a += w*c` style
 *     accumulations use fmaf explicitly (GPU compilers fuse them), everything else is
 *     separately rounded; mix(a,b,t) = a*(1-t) + b*t; length() = sqrtf(x*x + y*y) with the
 *     sum formed as fmaf(x, x, y*y); division by a shader constant is a true division
 *     unless stated otherwise;
 *   - transcendental functions are DEFINED, not libm's: the transfer curves of
 *     the hot path are the tabulated cubic composites of pl_oracle_curve.h (byte-identical
 *     generator in the product, csrc/curves.c), atan2 / sincos the fp32 polynomials of
 *     pl_oracle_math.h (pf_atan2f, pf_sincosf), everything else the binary64 algorithms of
 *     pl_oracle_math.h; orc_set_exact_math(1) routes the curves through the latter as well;
 *   - textures hold already-decoded float texel values (unorm: v/(2^n-1) by IEEE division;
 *     fp16: exact widening); nearest sampling picks floor(u*size) under the address mode,
 *     linear sampling is an fp32 mix of the four texels around (u*size - 0.5) — except at texel
 *     centres (1:1 direct sampling of a whole texture, taps of a separable pass whose other axis
 *     maps 1:1), where it is the texel itself, as on texture hardware;
 *   - linear 1-D LUTs follow the array form of src/shaders/lut.c:700-715.
 */
#ifndef PL_ORACLE_H_
#define PL_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_API __attribute__((visibility("default")))

// How the sampling position `pos` of output pixel gid is formed from the source rect
enum orc_coord_mode {
    // fp32 restatement of the compute-shader attribute emulation in
    // src/dispatch.c:1038-1062 (out_scale * (gid+0.5), nested mix of corner values)
    ORC_COORD_LITERAL = 0,
    // the same affine map evaluated in double and rounded once (what a fragment
    // shader's hardware interpolators deliver to within their sub-pixel precision)
    ORC_COORD_EXACT = 1,
};

// A decoded texture: w*h texels of `comps` interleaved floats
struct orc_tex {
    int w, h, comps;
    const float *data;
    int addr;   // sampler address mode: 0 clamp to edge, 1 repeat, 2 mirrored repeat
};

struct orc_rect { float x0, y0, x1, y1; };

// --- samplers (src/shaders/sampling.c) ---------------------------------------------

struct orc_polar_params {
    const float *lut;     // pl_filter->weights, src/filters.c:216-221
    int lut_entries;      // 256 (sampling.c:566)
    float radius;         // pl_filter->radius (after cutoff)
    float radius_zero;    // pl_filter->radius_zero (anti-ringing radius)
    float antiring;       // cfg.antiring
    float scale;          // setup_src `scale`
    uint8_t comp_mask;    // components to sample
    int coord_mode;       // enum orc_coord_mode
};

// pl_shader_sample_polar, sampling.c:587-912 (compute variant :723-783). dst = dw*dh vec4
ORC_API void orc_sample_polar(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                              const struct orc_polar_params *p, float *dst);

// built-in fast scalers (sampling.c:319-478); kind 0 bicubic, 1 hermite, 2 gaussian, 3 oversample
ORC_API void orc_sample_fast(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                             int kind, float scale, float rx, float ry, float threshold,
                             int coord_mode, float *dst);

struct orc_ortho_params {
    const float *lut;     // filter weights, 256 rows of row_stride floats (after fill_ortho_lut)
    int lut_entries;
    int row_size, row_stride;
    int use_linear;       // radius == radius_zero: linear-sampling trick (sampling.c:919-936)
    float antiring;
    float ratio;          // scaling ratio of this pass (anti-ringing only when > 1)
    float scale;
    uint8_t comp_mask;
    int pass;             // 0 = vertical, 1 = horizontal (SEP_VERT / SEP_HORIZ)
    int coord_mode;
};

// pl_shader_sample_ortho2, sampling.c:950-1104: one 1-D pass. dst = dw*dh vec4
ORC_API void orc_sample_ortho(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                              const struct orc_ortho_params *p, float *dst);

// pl_shader_sample_direct / nearest / bilinear (sampling.c:277-340): mode 0 nearest, 1 linear
ORC_API void orc_sample_direct(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                               int linear, float scale, int coord_mode, float *dst);

struct orc_deband_params {
    int iterations;
    float threshold, radius, grain;
    float grain_neutral[3];
    float scale;          // setup_src scale (pl_color_repr_normalize)
    uint8_t comp_mask;
    uint32_t frame_index; // SH_PARAMS(sh).index (uint8 counter of pl_dispatch_reset_frame)
};

// pl_shader_deband, sampling.c:183-275. Output size == src size, rect = full.
// (fx0, fy0) is the gl_FragCoord origin of texel (0,0): 0 when the deband shader is
// dispatched on its own, or the position of the plane inside a larger merged pass.
ORC_API void orc_deband(const struct orc_tex *src, const struct orc_deband_params *p, float *dst);

// sh_prng, src/shaders.c:965-998: advances state, writes 3 floats in [0,1]
ORC_API void orc_prng(uint32_t state[3], float out[3]);

// --- FBO / target storage rounding ---------------------------------------------------
// In-place round of n floats as if stored to and re-read from a texture of the format
// kind: 0 = fp16 (rgba16hf FBOs, renderer.c:395-397), 1 = unorm with `bits` bits, 2 = fp32
ORC_API void orc_store_round(float *data, size_t n, int kind, int bits);
// Quantise to integer code values (what lands in a unorm target): round(clamp(v)*max)
ORC_API void orc_store_unorm(const float *data, size_t n, int bits, uint16_t *out);
ORC_API void orc_store_half(const float *data, size_t n, uint16_t *out);

// --- colour (src/shaders/colorspace.c) -------------------------------------------------

// color.rgb = M*color.rgb + c  (colorspace.c:308, :569); m row-major
ORC_API void orc_apply_matrix(float *rgba, size_t npix, const float m[9], const float c[3]);
ORC_API void orc_scale(float *rgba, size_t npix, float scale, int with_alpha);

// pl_shader_linearize / pl_shader_delinearize, colorspace.c:589-847. csp_min/csp_max are
// the PL_HDR_NORM nominal luma of the colour space (pl_color_space_nominal_luma_ex)
ORC_API void orc_linearize(float *rgba, size_t npix, int trc, float csp_min, float csp_max,
                           const float luma_coeffs[3]);
ORC_API void orc_delinearize(float *rgba, size_t npix, int trc, float csp_min, float csp_max,
                             const float luma_coeffs[3], int black_scaled);

// BT.2100 ICtCp / BT.2020-CL post-processing of pl_shader_decode_color (:311-390) and the
// matching pre-processing of pl_shader_encode_color (:474-524). sys = enum pl_color_system
ORC_API void orc_decode_special(float *rgba, size_t npix, int sys);
ORC_API void orc_encode_special(float *rgba, size_t npix, int sys);

// pl_shader_sigmoidize / unsigmoidize, colorspace.c:851-894
ORC_API void orc_sigmoidize(float *rgba, size_t npix, float center, float slope);
ORC_API void orc_unsigmoidize(float *rgba, size_t npix, float center, float slope);

// pl_shader_dovi_reshape, colorspace.c:106-273. `dovi` = struct pl_dovi_metadata bytes
struct orc_dovi_comp {
    int num_pivots;
    float pivots[9];
    int method[8];
    float poly_coeffs[8][3];
    int mmr_order[8];
    float mmr_constant[8];
    float mmr_coeffs[8][3][7];
};
ORC_API void orc_dovi_reshape(float *rgba, size_t npix, const struct orc_dovi_comp comp[3]);
// the Dolby Vision tail of pl_shader_decode_color, colorspace.c:392-422: PQ EOTF, LMS -> RGB
// matrix (hard-coded HPE matrix x dovi->linear, row-major), PQ OETF
ORC_API void orc_dovi_lms(float *rgba, size_t npix, const float lms2rgb[9]);

// struct peak_buf_data, colorspace.c:936-942
#define ORC_PEAK_SLICES 12
#define ORC_PEAK_HIST_BINS 64
struct orc_peak_buf {
    uint32_t frame_wg_count[ORC_PEAK_SLICES];
    uint32_t frame_wg_active[ORC_PEAK_SLICES];
    uint32_t frame_sum_pq[ORC_PEAK_SLICES];
    uint32_t frame_max_pq[ORC_PEAK_SLICES];
    uint32_t frame_hist[ORC_PEAK_SLICES][ORC_PEAK_HIST_BINS];
};

struct orc_peak_params {
    int trc;                // source transfer
    float csp_min, csp_max; // nominal luma (PL_HDR_NORM) for linearize
    float luma_coeffs[3];   // RGB->Y row of rgb2xyz
    float black_cutoff;     // percent (pl_peak_detect_params.black_cutoff)
    int use_histogram;      // 0 < percentile < 100
    int wg_w, wg_h;         // work-group geometry of the (possibly merged) compute pass
};

// pl_shader_detect_peak GLSL, colorspace.c:1257-1350. Input = w*h vec4 image in the source
// colour space (non-linear); invocations beyond the image edge (when w,h are not multiples
// of the work-group size) replicate the clamped edge pixel like the reference's texture
// fetches do.
ORC_API void orc_detect_peak(const float *rgba, int w, int h, const struct orc_peak_params *p,
                             struct orc_peak_buf *out);

struct orc_peak_state { float avg_pq, max_pq; };
// update_peak_buf + measure_peak, colorspace.c:1020-1153. Returns 1 if data was consumed.
ORC_API int orc_peak_update(struct orc_peak_state *st, const struct orc_peak_buf *data,
                            float smoothing_period, float scene_low, float scene_high,
                            float percentile);

struct orc_color_map_params {
    int src_trc, dst_trc;
    float src_min, src_max, dst_min, dst_max;   // PL_HDR_NORM nominal luma for (de)linearize
    float src_luma[3], dst_luma[3];
    int dst_black_scaled;
    int prelinearized;
    // full IPT path (colorspace.c:1791-1995) when need_tone_map || need_gamut_map
    int need_tone_map, need_gamut_map;
    float rgb2lms[9], lms2rgb[9];   // already includes the saturation fast path product
    // tone map: kind 0 = LUT, 1 = clip (fast), 2 = linear (fast)
    int tone_kind;
    const float *tone_lut; int tone_lut_size;
    float tone_input_min, tone_input_max, tone_output_min, tone_output_max, tone_exposure;
    // gamut 3DLUT, rgba16 unorm texels as floats (v/65535), I fastest, then C, then h
    const float *gamut_lut; int lut_I, lut_C, lut_h;
    float gamut_min_luma, gamut_max_luma;
    // contrast recovery (:1879-1921): low-pass feature map (1 component), strength, and the
    // image geometry (pixels are row-major width x height) for the map's sampling position
    const struct orc_tex *feature_map;
    float contrast_recovery;
    int width, height;
};
// bilinear fetch of component 0 (textureLod(...).r) under the texture's address mode
float orc_tex_linear_r(const struct orc_tex *t, float u, float v);
ORC_API void orc_color_map(float *rgba, size_t npix, const struct orc_color_map_params *p);

// --- dithering (src/shaders/dithering.c) -----------------------------------------------
struct orc_dither_params {
    int method;          // enum pl_dither_method: 0 blue noise, 1 ordered LUT, 2 ordered fixed, 3 white
    const float *lut;    // lut_size*lut_size matrix
    int lut_size;        // in texels (1 << params->lut_size)
    int new_depth;
    int temporal;
    uint32_t frame_index;
    float gamma;         // approx_gamma(transfer), dithering.c:77-107
};
ORC_API void orc_dither(float *rgba, int w, int h, const struct orc_dither_params *p);

// pl_shader_error_diffusion, dithering.c:326-527. pattern = 3 rows x 5 cols (dy 0..2, dx -2..2)
ORC_API void orc_error_diffusion(const float *rgba_in, float *rgba_out, int w, int h,
                                 int new_depth, int shift, int divisor, const int *pattern);

// --- AV1 film grain (src/shaders/film_grain_av1.c) ---------------------------------------
struct orc_av1_grain_params {
    // LUT textures harvested from (or generated like) generate_grain_y/uv, :303-432
    const float *grain[3];     // per output comp: 64x64 (luma) or sub-sampled block
    int grain_w, grain_h;      // LUT dimensions of this plane's grain
    const uint32_t *offsets;   // packed block offsets, :434-462
    int offsets_w, offsets_h;
    const float *scaling[3];   // 256-entry scaling LUTs (NULL -> no scaling for that comp)
    int is_luma[3];            // channel is Y
    int num_comps;
    int sub_x, sub_y;          // plane subsampling relative to luma
    int overlap;
    int chroma_scaling_from_luma;
    int bit_depth;
    float scale;               // pl_color_repr_normalize
    float minValue, maxLuma, maxChroma; // clamp bounds in normalised units
    int scaling_shift;
    float uv_mult[3], uv_mult_luma[3], uv_offset[3];
    int luma_comp;
};
ORC_API void orc_av1_grain(const struct orc_tex *plane, const struct orc_tex *luma,
                           const struct orc_av1_grain_params *p, float *dst);

// pl_shader_extract_features (colorspace.c:1383-1404) on linear RGB; m = rgb2lms (float)
ORC_API void orc_extract_features(float *rgba, size_t npix, const float m[9]);
// pl_shader_deinterlace (shaders/deinterlacing.c): field 1 top / 2 bottom (0: passthrough), algo
// 0 weave, 1 bob, 2 yadif, 3 bwdif; prev / next may be NULL; dst: w*h*4 floats
ORC_API void orc_deinterlace(const struct orc_tex *cur, const struct orc_tex *prev,
                             const struct orc_tex *next, int field, int first_field, int algo,
                             int skip_spatial_check, unsigned comp_mask, float *dst);
// pl_render_image_mix accumulation (renderer.c:3955): acc = fma(w, rgba, acc), all four components
ORC_API void orc_mix_accumulate(float *acc, const float *rgba, size_t npix, float w);
// pl_shader_custom_lut (shaders/lut.c:212-280); data = RGB triples (red fastest), size[1] == 0: 1-D
ORC_API void orc_custom_lut(float *rgba, size_t npix, const float *data, const int size[3],
                            const float *shaper_in, const float *shaper_out);

// --- H.274 film grain (src/shaders/film_grain_h274.c) --------------------------------------
struct orc_h274_grain_params {
    const float *db;           // 832 x 832 grain data base (fill_grain_lut, :107-125)
    int db_w;
    uint32_t seed;
    float scale;               // pl_color_repr_normalize
    float scale_factor;        // 1 / (1 << (log2_scale_factor + 6))
    int num_entries;           // texture components with component_model_present, in order
    int channel[3];            // channel_map(idx)
    int num_intervals[3];
    const float *lower[3], *upper[3];   // bounds / 255
    const uint32_t *value[3];           // scale << 16 | h << 8 | v  (:276-282)
};
ORC_API void orc_h274_grain(const struct orc_tex *plane, const struct orc_h274_grain_params *p,
                            float *dst);

// --- double-precision ideal resampler (config C1) ----------------------------------------
// weight callback evaluates pl_filter_sample for distance x (from the reference via harvest)
typedef double (*orc_weight_fn)(void *priv, double x);
ORC_API void orc_polar_ideal(const struct orc_tex *src, int dw, int dh, double radius,
                             orc_weight_fn fn, void *priv, double *dst);

ORC_API int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif

// Array front-end to the portable math definitions of pl_oracle_math.h:
// fn 0 pow(a,b), 1 exp(a), 2 log(a), 3 sin(a), 4 cos(a), 5 atan2(a,b) (binary64 definitions),
// 6 sin, 7 cos, 8 atan2 (fp32 hot-path definitions), 9 log2(a), 16 + id: transfer curve `id`
// (pl_oracle_curve.h) with its exact fallback
void orc_math(int fn, const float *a, const float *b, float *out, size_t n);
ORC_API void orc_curve_apply(int id, const float *x, float *out, size_t n);
ORC_API void orc_set_exact_math(int on);
ORC_API int orc_curve_table(int id, int *base, int *nseg, int *mbits, float *coef);

// Evaluate ew x eh invocations (>= the output size) in orc_sample_polar / orc_sample_direct /
// orc_deband: the overhang of the last work groups, which peak detection counts. 0,0 = off.
ORC_API void orc_set_eval_extent(int ew, int eh);
