// Internal definitions shared by the host side and the CUDA kernels of libplacebo_b200.
// Nothing here is part of the C ABI (include/libplacebo_b200.h is).
#ifndef PLB_INTERNAL_H_
#define PLB_INTERNAL_H_

#include "../../include/libplacebo_b200.h"
#include "curves.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

// ---------------------------------------------------------------------------------------
// logging (reference: src/log.h PL_ERR/PL_WARN/... macros)
// ---------------------------------------------------------------------------------------
void plb_msg(pl_log log, enum pl_log_level lev, const char *fmt, ...)
    __attribute__((format(printf, 3, 4)));
#define PLB_ERR(log, ...)   plb_msg(log, PL_LOG_ERR, __VA_ARGS__)
#define PLB_WARN(log, ...)  plb_msg(log, PL_LOG_WARN, __VA_ARGS__)
#define PLB_INFO(log, ...)  plb_msg(log, PL_LOG_INFO, __VA_ARGS__)
#define PLB_DEBUG(log, ...) plb_msg(log, PL_LOG_DEBUG, __VA_ARGS__)

#define PLB_MIN(a, b) ((a) < (b) ? (a) : (b))
#define PLB_MAX(a, b) ((a) > (b) ? (a) : (b))
#define PLB_CLAMP(x, lo, hi) ((x) < (lo) ? (lo) : (x) > (hi) ? (hi) : (x))
#define PLB_DEF(x, d) ((x) ? (x) : (d))
#define PLB_ARRAY_SIZE(a) (sizeof(a) / sizeof((a)[0]))

// ---------------------------------------------------------------------------------------
// Texel formats as the kernels see them
// ---------------------------------------------------------------------------------------
enum plb_fmt_kind {
    PLB_FK_UNORM8 = 0,
    PLB_FK_UNORM16,
    PLB_FK_HALF,
    PLB_FK_FLOAT,
    PLB_FK_UINT32,
};

// Device-side view of a 2D (or 3D, for LUTs) texture: pitched linear memory in HBM
struct plb_tex_desc {
    void *ptr;
    size_t pitch;      // bytes per row
    int w, h;
    int comps;         // 1..4
    int kind;          // enum plb_fmt_kind
    // planar pair view (read only): component 0 from `ptr`, component 1 from `ptr2`, both
    // single-component planes of the same size and kind. Lets two chroma planes act as the
    // 2-component texture the reference would first merge them into (renderer.c:1617-1699),
    // without the copy pass: the values are identical.
    void *ptr2;
    size_t pitch2;
    // how fetches outside the texture resolve: enum pl_tex_address_mode of the sampling site
    // (pl_sample_src.address_mode); 0 = clamp to edge
    int addr;
};

// The private object behind a pl_tex (public part first, like PL_PRIV in src/pl_alloc.h)
struct plb_tex {
    struct pl_tex_t pub;
    pl_gpu gpu;
    struct plb_tex_desc d;
    bool owned;        // false for pl_cuda_wrap
    size_t alloc_size;
    // cross-stream ordering (async copy engines): sequence number of the last main-stream
    // operation that touched the texture, and the completion event of the last transfer
    uint64_t last_use;
    void *ev_xfer;     // cudaEvent_t, created on first transfer
    bool xfer_pending; // main stream has not yet waited for ev_xfer
    pl_tex parents[2]; // planar pair view: the textures it aliases
};
// Read-only 2-component view over two single-component textures (see plb_tex_desc.ptr2)
pl_tex plb_tex_view2(pl_gpu gpu, pl_tex a, pl_tex b, pl_fmt fmt);
static inline const struct plb_tex *plb_tex_priv(pl_tex tex) { return (const struct plb_tex *) tex; }

struct plb_buf {
    struct pl_buf_t pub;
    void *dptr;        // device memory (NULL for pure host buffers)
    void *hptr;        // pinned host mirror for host_readable / host_mapped buffers
};

struct plb_fmt {
    struct pl_fmt_t pub;
    int kind;
};

struct plb_gpu {
    struct pl_gpu_t pub;
    int device;
    void *stream;          // cudaStream_t
    bool own_stream;
    bool failed;
    bool no_timers;
    bool bank_ref;         // holds a reference on the device's transfer-curve tables
    bool exact_math;       // PLB_EXACT_MATH=1 at creation: specialised (curve-only) kernels are off
    int coord_mode;        // enum plb_coord_mode used by samplers
    uint64_t launches;     // kernels launched (pl_cuda_launch_count)
    int sm_count;
    struct plb_fmt *fmts;
    pl_fmt *fmt_list;
    void *staging;         // pinned staging buffer for pl_tex_upload/download
    size_t staging_size;
    // ring of small pinned staging slots for LUT-sized uploads (no stream synchronisation per
    // upload: a slot is reused only after the event recorded behind its copy kernel)
    void *up_ring;         // PLB_UP_SLOTS x PLB_UP_SLOT_BYTES, pinned
    void *up_event[8];
    int up_next;
    // Asynchronous transfers: uploads and downloads run on their own streams (the two copy
    // engines) and are ordered against the kernels by events, so that frame N+1 uploads and
    // frame N-1 downloads while frame N renders.
    bool async_xfer;
    void *h2d_stream, *d2h_stream;
    uint64_t seq;          // main-stream operations that touched textures
    void *use_ring[64];    // cudaEvent_t recorded after operation seq (slot seq % 64)
};
// Main-stream bookkeeping around every operation that reads or writes textures
void plb_gpu_before_use(pl_gpu gpu, pl_tex const *texs, int n);
void plb_gpu_after_use(pl_gpu gpu, pl_tex const *texs, int n);
static inline struct plb_gpu *plb_gpu_priv(pl_gpu gpu) { return (struct plb_gpu *) gpu; }
bool plb_cuda_check(pl_gpu gpu, int err, const char *what);
void *plb_gpu_staging(pl_gpu gpu, size_t size);

// small device allocations for LUTs / parameter blocks, owned by shader objects
void *plb_dev_alloc(pl_gpu gpu, size_t size);
void plb_dev_free(pl_gpu gpu, void *ptr);
bool plb_dev_upload(pl_gpu gpu, void *dst, const void *src, size_t size);
// byte-reverse every `wordsize`-byte word (2 or 4) of src[src_offset ..) into dst (src/gpu.h
// pl_buf_copy_swap); offsets and size are multiples of 4
bool plb_buf_copy_swap(pl_gpu gpu, pl_buf dst, size_t dst_offset, pl_buf src, size_t src_offset,
                       size_t size, int wordsize);
bool plb_dev_download(pl_gpu gpu, void *dst, const void *src, size_t size);
bool plb_dev_memset(pl_gpu gpu, void *dst, int val, size_t size);
bool plb_dev_copy(pl_gpu gpu, void *dst, const void *src, size_t size);

// ---------------------------------------------------------------------------------------
// Colour op program: the typed replacement of the GLSL text a pl_shader accumulates.
// Executed per pixel by plb_run_color_ops() (csrc/color_ops.cuh) inside the pixel, polar
// and ortho kernels, in recording order.
// ---------------------------------------------------------------------------------------
enum plb_op_kind {
    PLB_OP_NONE = 0,
    PLB_OP_SCALE,          // color *= f (optionally incl. alpha)
    PLB_OP_MATRIX,         // color.rgb = M*color.rgb + c
    PLB_OP_LINEARIZE,      // pl_shader_linearize
    PLB_OP_DELINEARIZE,    // pl_shader_delinearize
    PLB_OP_DECODE_SPECIAL, // ICtCp / BT.2020-CL tail of pl_shader_decode_color
    PLB_OP_ENCODE_SPECIAL,
    PLB_OP_SIGMOID,
    PLB_OP_UNSIGMOID,
    PLB_OP_COLOR_MAP,      // IPT tone/gamut mapping core of pl_shader_color_map_ex
    PLB_OP_DITHER,         // pl_shader_dither
    PLB_OP_ROUND_FBO,      // emulate a store+load through an intermediate FBO format
    PLB_OP_ALPHA,          // pl_shader_set_alpha conversions
    PLB_OP_DOVI_RESHAPE,   // pl_shader_dovi_reshape
    PLB_OP_DOVI_LMS,       // Dolby Vision tail of decode_color: PQ EOTF, LMS->RGB, PQ OETF
    PLB_OP_GAMMA,          // color adjustment gamma
    PLB_OP_PEAK,           // pl_shader_detect_peak tap (needs block-wide reduction)
    PLB_OP_SWIZZLE,        // swizzle_color (src/renderer.c:791-808)
    PLB_OP_FEATURES,       // pl_shader_extract_features: linear RGB -> IPT intensity
    PLB_OP_CUSTOM_LUT,     // pl_shader_custom_lut (1-D per channel, 3-D tetrahedral)
};

struct plb_trc_consts {
    int trc;
    float csp_min, csp_max;
    float luma[3];
    float a, b;            // BT.1886 a, b (or 1/a, b for delinearize); HLG: y, b
    int black_scaled;      // delinearize pre-scale enabled
};

struct plb_op_color_map {
    float rgb2lms[9], lms2rgb[9];
    int need_tone, need_gamut;
    int tone_kind;                     // 0 LUT, 1 clip, 2 linear
    const float *tone_lut; int tone_lut_size;
    float tone_a, tone_b;              // 1/range, -min/range
    float in_min, in_max, out_min, out_max;
    float lin_a, lin_b;                // linear fast path
    const float *gamut_lut;            // rgba32f texels (quantised to n/65535), I fastest
    int lut_I, lut_C, lut_h;
    float g_a, g_b;
    float g_nm1[3];                    // LUT addressing per axis: (float) (n - 1)
    // contrast recovery (colorspace.c:1879-1921): low-pass feature map (r16hf, PQ intensity),
    // bicubic through four bilinear fetches; positions relative to the pass's output size
    int has_fmap;
    struct plb_tex_desc fmap;
    float recovery;
    float fm_size[2], fm_pt[2];        // (float) size and (float) (1.0 / size) of the map
    int out_w, out_h;                  // size of the pass's output (set at launch)
};

struct plb_op_dither {
    int method;
    const float *lut; int lut_size;
    int new_depth;
    int temporal;
    uint32_t frame_index;
    float gamma;
    float rot[4];
    // derived on the host: power-of-two LUT, no temporal rotation, plain (not gamma-aware) rounding
    int fast_pow2;
    float scale, inv_scale;            // 2^depth - 1 and its reciprocal
};

struct plb_op_peak {
    struct plb_trc_consts lin;
    float cutoff;
    int use_histogram;
    uint32_t *buf;                     // device struct peak_buf_data (3264 B)
};

struct plb_dovi_dev;                   // uploaded reshape tables

struct plb_color_op {
    int kind;
    union {
        struct { float s; int with_alpha; } scale;
        struct { float m[9]; float c[3]; int has_c; } matrix;
        struct plb_trc_consts trc;
        struct { int sys; } special;
        struct { float center, slope, offset, scale; } sigmoid;
        struct plb_op_color_map cmap;
        struct plb_op_dither dither;
        struct { int kind; int bits; } round_fbo;
        struct { int from, to; } alpha;
        struct { int unused; } dovi;      // tables: plb_pixel_launch.dovi
        struct { float g; } gamma;
        struct plb_op_peak peak;
        struct { int8_t map[4]; int comps; int force_alpha; } swizzle;
        struct { float m[9]; } features;  // (SDR_WHITE / 10000) * rgb2lms
        struct {                          // pl_shader_custom_lut
            const float *lut;             // device, 4 floats per entry (r fastest)
            int size[3];                  // size[1] == 0: 1-D
            int has_in, has_out;
            float shaper_in[9], shaper_out[9];
        } custom_lut;
    } u;
};

#define PLB_MAX_OPS 12
struct plb_color_prog {
    int n;
    struct plb_color_op ops[PLB_MAX_OPS];
};

// PQ constants as the reference's GLSL text carries them ("%f" = 6 decimals), computed on
// the host once (csrc/shaders.cpp: plb_pq_literal) and handed to kernels in constant memory
struct plb_pq_consts {
    // transfer curves (curves.h), filled by gpu.cu after the tables are uploaded
    struct plb_curve cv[PLB_CV_COUNT];
    const void *cv_bank;   // all tables, curve id at entry plb_cv_offset(id) (float4 units)
    int exact_math;        // PLB_EXACT_MATH=1: round-1 binary64 definitions everywhere
    float m1, m2, c1, c2, c3, inv_m1, inv_m2, to_pq, from_pq, inv_from_pq;
    float hlg_a, hlg_b, hlg_c, hlg_inv_a;
    float inv_2pi;
    // camera log curves (colorspace.c:629-655, 789-812), constants as printed by "%f"
    float vlog_b, vlog_d, vlog_inv_c, vlog_c10;
    float slog_b, slog_c, slog_inv_a, slog_a10, slog_p, slog_inv_p, slog_q, slog_k2, slog_inv_k2;
};

// ---------------------------------------------------------------------------------------
// Sampling stage descriptions
// ---------------------------------------------------------------------------------------
enum plb_coord_mode { PLB_COORD_LITERAL = 0, PLB_COORD_EXACT = 1 };

// A sampler launch = primary sampler + the common per-pixel tail (struct plb_pixel_launch:
// extra planes to merge, colour program, target). Declared below.
struct plb_pixel_launch;
enum plb_plane_kind {
    PLB_PLANE_NONE = 0,
    PLB_PLANE_DIRECT,      // pl_shader_sample_direct / nearest / bilinear
    PLB_PLANE_DEBAND,      // pl_shader_deband
    PLB_PLANE_GRAIN_AV1,   // pl_shader_film_grain (AV1)
    PLB_PLANE_GRAIN_H274,
    PLB_PLANE_BICUBIC,     // pl_shader_sample_bicubic   (sampling.c:319-366)
    PLB_PLANE_HERMITE,     // pl_shader_sample_hermite   (sampling.c:368-393)
    PLB_PLANE_GAUSSIAN,    // pl_shader_sample_gaussian  (sampling.c:395-440)
    PLB_PLANE_OVERSAMPLE,  // pl_shader_sample_oversample (sampling.c:442-478)
};

struct plb_av1_dev;        // uploaded grain LUTs, see k_grain.cu
struct plb_h274_dev;

// AV1 film grain for one plane pass (film_grain_av1.c:605-1013): device LUTs + constants.
// Component i of the plane's texture is colour channel `channel[i]` (PL_CHANNEL_Y/CB/CR or -1).
struct plb_av1_plane {
    int ncomp;
    int channel[3];
    const float *grain[3];     // planar 2-D grain LUT of component i (NULL: no grain)
    int grain_w[3];            // its row length (64 or 32)
    const float *scaling[3];   // 256-entry scaling LUT of component i (NULL: skip component)
    const uint32_t *offsets;   // packed block offsets (TL, T, L, N), one word per 32x32 block
    int off_w;
    int sub_x, sub_y, overlap;
    float tex_scale, grain_min, grain_max;
    float min_value, max_luma, max_chroma;
    int chroma_from_luma;
    float mult[3][2];          // (uv_mult_luma, uv_mult) / 64 for component i
    float offset[3];
    int luma_from_comp;        // >= 0: averageLuma = colour component of this texture
    int luma_comp;             // else component of the separate luma texture
};

// H.274 / SMPTE RDD 5 film grain for one plane pass (film_grain_h274.c:145-315). One device
// record per recorded pass: this header followed by the intensity intervals of all entries.
struct plb_h274_interval {
    float lo, hi;              // bounds / 255
    uint32_t value;            // scale << 16 | h << 8 | v
};
struct plb_h274_plane {
    const float *db;           // (13*64)^2 grain data base, row length 13*64
    int nent;                  // texture components that receive grain (in component order)
    uint32_t seed;
    float tex_scale, scale_factor;
    struct {
        int channel;           // PL_CHANNEL_*: indexes the colour vector and the pcg3d lanes
        int first, count;      // its intervals in iv[]
    } ent[3];
    struct plb_h274_interval iv[];
};

struct plb_plane_src {
    int kind;
    struct plb_tex_desc tex;
    pl_rect2df rect;       // source rect in texels, mapped over the output
    int linear;            // bilinear instead of nearest (DIRECT)
    float scale;
    int8_t map[4];         // colour index written by texture component k (-1 = none)
    // deband
    int iterations;
    float threshold, radius, grain_str;
    float neutral[3];
    uint8_t comp_mask;
    // oversample: scaling ratios and snapping threshold
    float over_rx, over_ry, over_threshold;
    // grain
    struct plb_av1_plane av1;
    struct plb_tex_desc luma;  // AV1 chroma needs the luma plane
    const struct plb_h274_plane *h274; // device copy (rarely used: kept out of line)
};

// Dolby Vision reshaping tables as the reference packs them for the GPU
// (shaders/colorspace.c:128-169): per component up to 8 pieces with (c0, c1, c2, order) or
// (constant, mmr_idx, -, order), the internal pivots padded with 1e9, packed MMR weights.
struct plb_dovi_tab {
    int num_pivots[3];
    float lo[3], hi[3];
    float pivots[3][7];
    float coeffs[3][8][4];
    float mmr[3][48][4];
};

struct plb_pixel_launch {
    int nplanes;
    struct plb_plane_src planes[4];
    float init[4];         // colour before planes are merged (renderer.c:1792)
    struct plb_tex_desc dst;   // dst.ptr == NULL: no store (peak measurement only)
    pl_rect2d dst_rect;
    int out_w, out_h;
    int coord_mode;
    uint32_t frame_index;
    int wg_w, wg_h;        // work-group geometry (matters to PLB_OP_PEAK only)
    int has_dovi;
    struct plb_dovi_tab dovi;  // operand of PLB_OP_DOVI_RESHAPE (at most one per pass)
    struct plb_color_prog prog;
};
bool plb_launch_pixel(pl_gpu gpu, const struct plb_pixel_launch *p);

struct plb_polar_launch {
    struct plb_tex_desc src;
    pl_rect2df rect;           // source rect (texels)
    const float *lut;          // device copy of pl_filter->weights (256 floats)
    const float *host_lut;     // host copy (phase tables of the 2x fast path)
    int lut_entries;
    float radius, radius_zero, antiring, scale;
    uint8_t comp_mask;
    int allow_fast;            // phase-table fast path permitted (needs PLB_COORD_EXACT)
    int8_t src_map[4];         // colour index receiving sampled component k
    struct plb_pixel_launch tail;
};
bool plb_launch_polar(pl_gpu gpu, const struct plb_polar_launch *p);

struct plb_ortho_launch {
    struct plb_tex_desc src;
    pl_rect2df rect;
    const float *lut;          // device: 256 rows * row_stride (after fill_ortho_lut)
    int lut_entries, row_size, row_stride;
    int use_linear;
    float antiring, ratio, scale;
    uint8_t comp_mask;
    int pass;                  // 0 vertical, 1 horizontal
    int8_t src_map[4];
    struct plb_pixel_launch tail;
};
bool plb_launch_ortho(pl_gpu gpu, const struct plb_ortho_launch *p);

// frame mixing (pl_render_image_mix, renderer.c:3906-3966): weighted sum of up to 8 same-sized
// frames, each linearised first when more than one takes part
#define PLB_MAX_MIX 8
struct plb_mix_launch {
    int n;
    struct plb_tex_desc tex[PLB_MAX_MIX];
    float weight[PLB_MAX_MIX];     // already divided by the weight sum
    int linearize;                 // frames are taken to linear light before mixing
    struct plb_trc_consts lin;
    int8_t src_map[4];
    struct plb_pixel_launch tail;
};
bool plb_launch_mix(pl_gpu gpu, const struct plb_mix_launch *p);

// pl_shader_deinterlace (shaders/deinterlacing.c)
struct plb_deint_launch {
    struct plb_tex_desc cur, prev, next;   // prev / next = cur when absent
    int field;                             // enum pl_field
    int prev2_is_prev, next2_is_cur;
    int algo, skip_spatial_check, intra_only;
    uint8_t comp_mask;
    int8_t src_map[4];
    struct plb_pixel_launch tail;
};
bool plb_launch_deint(pl_gpu gpu, const struct plb_deint_launch *p);

struct plb_errdiff_launch {
    struct plb_tex_desc src, dst;
    int w, h;
    int new_depth;
    int shift, divisor;
    int pattern[3][5];
};
bool plb_launch_errdiff(pl_gpu gpu, const struct plb_errdiff_launch *p);

void plb_pq_literal(struct plb_pq_consts *q);
bool plb_curve_generate(int id, const struct plb_pq_consts *q, struct plb_curve_host *out);
bool plb_upload_pq_consts(pl_gpu gpu);
bool plb_gpu_no_timers(pl_gpu gpu);

#ifdef __cplusplus
}
#endif
#endif
