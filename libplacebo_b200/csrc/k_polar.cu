// Polar (EWA) resampling — the CUDA replacement of the GLSL emitted by
// pl_shader_sample_polar (src/shaders/sampling.c:587-912, per-tap body polar_sample :503-558).
//
// Two kernels:
//
//  k_polar_generic   any ratio / rect / filter / anti-ringing. One thread per output pixel,
//                    per-pixel sub-texel fraction and per-tap weight (sqrt + LUT lerp), source
//                    texels straight from L1/L2. Follows the oracle operation by operation in
//                    both coordinate modes, so it is bit-exact with it up to store rounding.
//
//  k_polar_up2<NC,R2> exact 2x upscale of a whole texture (chroma 4:2:0 -> luma size, 1080p ->
//                    4K): with exact coordinates the sub-texel fraction takes only the values
//                    {0.75, 0.25} per axis, so the 4 x 64 tap weights are tabulated once on the
//                    host (same fp32 ops as the per-pixel path -> identical values) and handed
//                    to the kernel as __grid_constant__ parameters (constant bank operands of
//                    the FFMAs). Each thread produces a 4x2 output block from a 10x9 source
//                    window held in registers, read as float2 from a planar fp32 tile staged in
//                    shared memory; taps are visited in the reference's (y, x) order and
//                    accumulated with FFMA, so results equal the exact-coordinate oracle bit
//                    for bit. Taps that can never contribute for radius^2 < R2/16 are removed
//                    at compile time (30 of 64 for EWA-Lanczos).
//
// Roofline notes: 2x EWA-Lanczos needs 30 taps x NC FFMA per output
// pixel; at 128 FFMA/clk/SM that is the binding limit (not HBM): 0.70 clk/px/SM for NC=3.
#include "pixel_core.cuh"

#include <cuda.h> // CUtensorMap (driver API types; the entry point is resolved at run time)

void plb_fill_tail(plb_tail_args *t, const struct plb_pixel_launch *p);

PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_polar)

#define MAX_TAPS 144

struct polar_generic_args {
    plb_tex_desc src;
    plb_coord cc;
    const float *lut;
    int lut_entries;
    float radius, inv_radius, radius_zero, antiring, scale;
    uint8_t comp_mask;
    int ntaps;
    int8_t tap_x[MAX_TAPS], tap_y[MAX_TAPS];
    uint8_t tap_flags[MAX_TAPS]; // bit0 guard, bit1 anti-ringing candidate
    plb_tail_args tail;
};

// the EWA sum for one output pixel; out = vec4 as pl_shader_sample_polar leaves `color`
PLB_DEV void polar_pixel(const polar_generic_args &a, int gx, int gy, float out[4])
{
    int jx, jy;
    float fx, fy;
    plb_coord_texel(a.cc, gx, gy, jx, jy, fx, fy);
    float color[4] = {0, 0, 0, 0};
    float wsum = 0.0f;
    float ar[4][2] = {{0}}, wwsum[4][2] = {{0}};
    const bool use_ar = a.antiring > 0.0f;

    for (int t = 0; t < a.ntaps; t++) {
        const float ox = (float) a.tap_x[t], oy = (float) a.tap_y[t];
        const float dx = ox - fx, dy = oy - fy;
        const float d = sqrtf(__fmaf_rn(dx, dx, dy * dy));
        if ((a.tap_flags[t] & 1) && !(d < a.radius))
            continue;
        const float w = plb_lut1d(a.lut, a.lut_entries, d * a.inv_radius);
        wsum += w;
        const float4 c4 = plb_load_clamped(a.src, jx + a.tap_x[t], jy + a.tap_y[t]);
        const float c[4] = {c4.x, c4.y, c4.z, c4.w};
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!(a.comp_mask & (1 << k)) || k >= a.src.comps)
                continue;
            color[k] = __fmaf_rn(w, c[k], color[k]);
            if ((a.tap_flags[t] & 2) && d <= a.radius_zero) {
                float cc0 = 1.0f - a.scale * c[k], cc1 = a.scale * c[k];
                float ww0 = cc0 + 0.10f, ww1 = cc1 + 0.10f;
                #pragma unroll
                for (int i = 0; i < 5; i++) {
                    ww0 = ww0 * ww0;
                    ww1 = ww1 * ww1;
                }
                ww0 = w * ww0;
                ww1 = w * ww1;
                ar[k][0] = __fmaf_rn(ww0, cc0, ar[k][0]);
                ar[k][1] = __fmaf_rn(ww1, cc1, ar[k][1]);
                wwsum[k][0] += ww0;
                wwsum[k][1] += ww1;
            }
        }
    }

    const float norm = a.scale / wsum;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool have = (a.comp_mask & (1 << k)) && k < a.src.comps;
        float v = norm * color[k];
        if (have && use_ar) {
            float wx = ar[k][0] / wwsum[k][0], wy = ar[k][1] / wwsum[k][1];
            wx = 1.0f - wx;
            float w = plb_clamp(v, wx, wy);
            if (wx > wy)
                w = wx * 0.5f + wy * 0.5f;
            v = plb_mix(v, w, a.antiring);
        }
        out[k] = have ? v : 0.0f;
    }
    if (!(a.comp_mask & 8))
        out[3] = 1.0f;
}

// CTA = 32x8 threads over a 32x32 output tile, like k_pixel
__global__ void __launch_bounds__(256)
k_polar_generic(const __grid_constant__ polar_generic_args a)
{
    __shared__ plb_peak_smem ps;
    plm_tables_init();
    plb_tail_stage(a.tail);
    __syncthreads();
    const plb_tail_args &t = plb_s_tail;
    const int X0 = blockIdx.x * 32, Y0 = blockIdx.y * 32;
    const int gx = X0 + threadIdx.x;
    bool has_peak = false;
    for (int i = 0; i < plb_s_prog.n; i++)
        has_peak |= plb_s_prog.ops[i].kind == PLB_OP_PEAK;
    const int wgs_x = has_peak ? 32 / t.wg_w : 1, wgs_y = has_peak ? 32 / t.wg_h : 1;
    const int nwgx = has_peak ? (t.out_w + t.wg_w - 1) / t.wg_w : 1;
    const int nwgy = has_peak ? (t.out_h + t.wg_h - 1) / t.wg_h : 1;
    if (has_peak)
        plb_peak_begin(ps);
    const plb_op_peak *pk_any = nullptr;
    #pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const int gy = Y0 + threadIdx.y + 8 * r;
        const bool inside = gx < t.out_w && gy < t.out_h;
        if (!has_peak && !inside)
            continue;
        float prim[4], color[4];
        polar_pixel(a, gx, gy, prim);
        uint32_t ypq = 0;
        const plb_op_peak *pk = nullptr;
        plb_tail_color(gx, gy, prim, color, &ypq, &pk);
        if (has_peak) {
            const int lx = (gx - X0) / t.wg_w, ly = (gy - Y0) / t.wg_h;
            const bool valid = (X0 / t.wg_w + lx) < nwgx && (Y0 / t.wg_h + ly) < nwgy;
            pk_any = pk;
            plb_peak_add(ps, *pk, valid, ly * wgs_x + lx, ypq);
        }
        if (inside && t.dst.ptr)
            plb_store(t.dst, t.dst_x0 + gx, t.dst_y0 + gy,
                      make_float4(color[0], color[1], color[2], color[3]));
    }
    if (has_peak)
        plb_peak_finish(ps, *pk_any, wgs_x, wgs_y, X0 / t.wg_w, Y0 / t.wg_h, nwgx, nwgy,
                        (uint32_t) (t.wg_w * t.wg_h));
}

// ---------------------------------------------------------------------------------------
// 2x fast path
// ---------------------------------------------------------------------------------------

#define UP2_TX 16           // threads per CTA in x
#define UP2_TY 16
#define UP2_OX 8            // outputs per thread in x (four source cells)
#define UP2_OY 2
#define UP2_TW (UP2_TX * UP2_OX)               // 128 output columns per CTA
#define UP2_TH (UP2_TY * UP2_OY)               // 32 output rows
#define UP2_TILE_W (UP2_TW / 2 + 8)            // 72 source columns (multiple of 4: float4 reads)
#define UP2_TILE_H (UP2_TH / 2 + 8)            // 24 source rows

struct polar_up2_args {
    plb_tex_desc src;
    float w[2][2][8][8];    // [phase_y][phase_x][oy+3][ox+3]; phase 0 = fraction 0.75
    float norm[2][2];       // scale / wsum per phase
    int use_tma;            // interior tiles come in through the tensor map
    int src_kind;           // UP2_SRC_*: source layout with a dedicated tile loader
    plb_tail_args tail;
};

// Sub-texel phases of an exact 2x upscale, per axis. MODE 0 ("centred": rect origin 0, e.g.
// RGB frames or centre-sited chroma): fractions 3/4, 1/4 and the first output of a cell starts
// one texel to the left. MODE 1 ("left"/"top" sited chroma, rect origin 0.25): fractions 0, 1/2,
// both outputs share the cell's texel as base.
__host__ __device__ constexpr int up2_frac4(int mode, int phase)
{
    return mode == 0 ? (phase ? 1 : 3) : (phase ? 2 : 0);
}
__host__ __device__ constexpr int up2_shift(int mode, int phase)
{
    return mode == 0 && phase == 0 ? 1 : 0;
}

// Is tap (ox, oy) of phase (py, px) inside radius^2 < R2_16/16 ?
template <int R2_16, int XM, int YM>
__host__ __device__ constexpr bool up2_tap_on(int py, int px, int oy, int ox)
{
    return ox >= -3 && ox <= 4 && oy >= -3 && oy <= 4 &&
           ((4 * ox - up2_frac4(XM, px)) * (4 * ox - up2_frac4(XM, px)) +
            (4 * oy - up2_frac4(YM, py)) * (4 * oy - up2_frac4(YM, py))) < R2_16;
}

// ---- TMA (cp.async.bulk.tensor) staging of the raw source tile -------------------------------
// The tile of an interior CTA is one 2-D box of the source texture: a single thread issues the
// bulk copy, the copy engine writes shared memory and signals an mbarrier with the byte count;
// nobody computes addresses or waits on individual loads. Tiles that touch the texture border
// need clamp-to-edge (TMA fills out-of-range texels with zeros), they take the manual loader.
PLB_DEV uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

PLB_DEV void tma_load_tile(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar, uint32_t bytes)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(b) : "memory");
}

PLB_DEV void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(b), "r"(parity) : "memory");
}

// Source layouts with a dedicated loader (launch-uniform; 0 = any, through plb_load_clamped)
enum { UP2_SRC_ANY = 0, UP2_SRC_HALF4 = 1, UP2_SRC_HALF2 = 2 };

template <int NC, int MODE>
constexpr int up2_smem_bytes()
{
    const int tile = NC * UP2_TILE_H * UP2_TILE_W * 4;
    const int raw = UP2_TILE_H * UP2_TILE_W * 8; // raw TMA tile (largest texel: rgba16hf)
    const int stage = MODE == 2 ? NC * UP2_TH * UP2_TW * 4 : 0;
    return tile + (raw > stage ? raw : stage) + 16;
}

// PURE: a resampling pass with nothing to merge and at most a light colour program (config
// C2's main scaler: 1/scale + swizzle): register-resident tail and vector stores.
//
// MODE 0: general tail (any planes, any colour program; plb_tail_color, not inlined).
// MODE 1 = PURE (above). MODE 2 = the merged plane pass of an HDR render: directly sampled or
// debanded / AV1-grained planes next to the EWA-scaled ones, light steps and the peak-detection
// tap with PQ input. The normalised filter results go through shared memory so that (a) ONE
// inlined copy of the tail serves the thread's sixteen outputs (the accumulators need
// compile-time indices, the staged values do not) and (b) the tail runs x-fastest: a warp
// handles 32 consecutive pixels of a row, so plane fetches and stores are contiguous. Per-pixel
// arithmetic is that of MODE 0.
//
// Thread = 8 x 2 outputs (four source cells) from a 12 x 9 source window: per window row three
// LDS.128 per component, then the row's taps as FFMAs whose weights are constant-bank operands.
// 48 accumulators (NC = 3) give the FMA pipe independent work without many resident warps.
template <int NC, int XM, int YM, int MODE>
__global__ void __launch_bounds__(UP2_TX * UP2_TY, 2)
k_polar_up2(const __grid_constant__ polar_up2_args a, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int R2_16 = 160; // taps beyond radius^2 = 10 never contribute (host check)
    constexpr bool PURE = MODE == 1;
    constexpr int TW = UP2_TW, TH = UP2_TH;
    constexpr int TILE_BYTES = NC * UP2_TILE_H * UP2_TILE_W * 4;
    const int SRC = a.src_kind; // launch-uniform
    // the raw TMA tile is dead once converted; the MODE 2 staging buffer reuses its space
    extern __shared__ __align__(128) unsigned char smem[]; // up2_smem_bytes<NC, MODE>()
    __shared__ __align__(8) uint64_t bar;
    float (*tile)[UP2_TILE_H][UP2_TILE_W] = (float (*)[UP2_TILE_H][UP2_TILE_W]) smem;
    unsigned char *aux = smem + TILE_BYTES;

    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * UP2_TX + tx;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int col0 = X0 / 2 - 4, row0 = Y0 / 2 - 4;

    // PURE passes never touch the transcendental tables and read the launch description
    // straight from the kernel parameters (constant bank): nothing to stage but the tile.
    if (!PURE) {
        plm_tables_init(); // published by the barrier after the tile staging
        if (MODE != 2)
            plb_tail_stage(a.tail);
    }

    const bool interior = SRC != UP2_SRC_ANY && a.use_tma && col0 >= 0 && row0 >= 0 &&
                          col0 + UP2_TILE_W <= a.src.w && row0 + UP2_TILE_H <= a.src.h;
    if (interior) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            tma_load_tile(aux, &tmap, col0, row0, &bar,
                          UP2_TILE_H * UP2_TILE_W * (SRC == UP2_SRC_HALF4 ? 8 : 4));
        }
        __syncthreads(); // barrier initialised before anyone polls it
        mbar_wait(&bar, 0);
        for (int i = tid; i < UP2_TILE_H * UP2_TILE_W; i += UP2_TX * UP2_TY) {
            const int r = i / UP2_TILE_W, c = i - r * UP2_TILE_W;
            if (SRC == UP2_SRC_HALF4) {
                const uint2 raw = ((const uint2 *) aux)[i];
                const float2 lo = __half22float2(*(const __half2 *) &raw.x);
                const float2 hi = __half22float2(*(const __half2 *) &raw.y);
                tile[0][r][c] = lo.x;
                if (NC > 1) tile[1 % NC][r][c] = lo.y;
                if (NC > 2) tile[2 % NC][r][c] = hi.x;
                if (NC > 3) tile[3 % NC][r][c] = hi.y;
            } else {
                const float2 v = __half22float2(((const __half2 *) aux)[i]);
                tile[0][r][c] = v.x;
                if (NC > 1) tile[1 % NC][r][c] = v.y;
            }
        }
    } else {
        // clamp-to-edge staging (border tiles, other layouts), planar fp32
        for (int i = tid; i < UP2_TILE_H * UP2_TILE_W; i += UP2_TX * UP2_TY) {
            const int r = i / UP2_TILE_W, c = i - r * UP2_TILE_W;
            float4 v;
            if (SRC == UP2_SRC_HALF4) {
                const int x = plb_clampi(col0 + c, 0, a.src.w - 1), y = plb_clampi(row0 + r, 0, a.src.h - 1);
                const uint2 raw = *(const uint2 *) ((const char *) a.src.ptr + (size_t) y * a.src.pitch + (size_t) x * 8);
                const float2 lo = __half22float2(*(const __half2 *) &raw.x);
                const float2 hi = __half22float2(*(const __half2 *) &raw.y);
                v = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else if (SRC == UP2_SRC_HALF2) {
                const int x = plb_clampi(col0 + c, 0, a.src.w - 1), y = plb_clampi(row0 + r, 0, a.src.h - 1);
                const float2 f = __half22float2(*(const __half2 *) ((const char *) a.src.ptr + (size_t) y * a.src.pitch + (size_t) x * 4));
                v = make_float4(f.x, f.y, 0.0f, 1.0f);
            } else {
                v = plb_load_clamped(a.src, col0 + c, row0 + r);
            }
            tile[0][r][c] = v.x;
            if (NC > 1) tile[1 % NC][r][c] = v.y;
            if (NC > 2) tile[2 % NC][r][c] = v.z;
            if (NC > 3) tile[3 % NC][r][c] = v.w;
        }
    }
    __syncthreads();

    float acc[UP2_OY][UP2_OX][NC];
    #pragma unroll
    for (int i = 0; i < UP2_OY; i++)
        #pragma unroll
        for (int j = 0; j < UP2_OX; j++)
            #pragma unroll
            for (int k = 0; k < NC; k++)
                acc[i][j][k] = 0.0f;

    // thread's source cells: columns i0 = X0/2 + 4*tx (+0..3), row jj = Y0/2 + ty
    // window column wx <-> source column i0 - 4 + wx  (tile column 4*tx + wx)
    // window row    wy <-> source row    jj - 4 + wy  (tile row    ty + wy)
    #pragma unroll
    for (int wy = 0; wy < 9; wy++) {
        float v[NC][12];
        #pragma unroll
        for (int k = 0; k < NC; k++) {
            const float4 *rowp = (const float4 *) &tile[k][ty + wy][4 * tx];
            #pragma unroll
            for (int q = 0; q < 3; q++) {
                const float4 f = rowp[q];
                v[k][4 * q] = f.x;
                v[k][4 * q + 1] = f.y;
                v[k][4 * q + 2] = f.z;
                v[k][4 * q + 3] = f.w;
            }
        }
        #pragma unroll
        for (int oyo = 0; oyo < UP2_OY; oyo++) {
            // output row 2*jj + oyo: phase oyo, base row jj - 1 (phase 0) or jj (phase 1)
            const int oy = wy - 4 + up2_shift(YM, oyo);
            #pragma unroll
            for (int oxo = 0; oxo < UP2_OX; oxo++) {
                const int ci = oxo >> 1, px = oxo & 1;
                #pragma unroll
                for (int wx = 0; wx < 12; wx++) {
                    const int ox = wx - 4 - ci + up2_shift(XM, px);
                    if (up2_tap_on<R2_16, XM, YM>(oyo, px, oy, ox)) {
                        const float w = a.w[oyo][px][oy + 3][ox + 3];
                        #pragma unroll
                        for (int k = 0; k < NC; k++)
                            acc[oyo][oxo][k] = __fmaf_rn(w, v[k][wx], acc[oyo][oxo][k]);
                    }
                }
            }
        }
    }

    if (PURE) {
        const plb_tail_args &t = a.tail;
        // launch-uniform: the usual identity component mapping needs no per-pixel selects
        const bool ident = t.src_map[0] == 0 && (NC < 2 || t.src_map[1] == 1) &&
                           (NC < 3 || t.src_map[2] == 2) && (NC < 4 || t.src_map[3] == 3);
        const int gx = X0 + tx * UP2_OX;
        if (gx >= t.out_w)
            return;
        #pragma unroll
        for (int oyo = 0; oyo < UP2_OY; oyo++) {
            const int gy = Y0 + ty * UP2_OY + oyo;
            if (gy >= t.out_h)
                continue;
            #pragma unroll
            for (int half = 0; half < UP2_OX / 4; half++) {
                float col[4][4];
                #pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float norm = a.norm[oyo][p & 1];
                    float prim[4] = {0.0f, 0.0f, 0.0f, 1.0f};
                    #pragma unroll
                    for (int k = 0; k < NC; k++)
                        prim[k] = norm * acc[oyo][4 * half + p][k];
                    #pragma unroll
                    for (int j = 0; j < 4; j++)
                        col[p][j] = t.init[j];
                    if (ident) {
                        #pragma unroll
                        for (int k = 0; k < NC; k++)
                            col[p][k] = prim[k];
                    } else {
                        #pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int m = t.src_map[k];
                            #pragma unroll
                            for (int j = 0; j < 4; j++)
                                col[p][j] = m == j ? prim[k] : col[p][j];
                        }
                    }
                }
                const int gxh = gx + 4 * half;
                // directly sampled extra planes (e.g. the luma plane next to EWA-scaled chroma)
                #pragma unroll 1
                for (int i = 0; i < t.nplanes; i++) {
                    const plb_plane_src &pl = t.planes[i];
                    const plb_coord &cc = t.plane_cc[i];
                    const bool direct = cc.ident; // texel centres: bilinear == the texel
                    const float sc = pl.scale;
                    int8_t map[4];
                    #pragma unroll
                    for (int k = 0; k < 4; k++)
                        map[k] = pl.map[k];
                    #pragma unroll
                    for (int p = 0; p < 4; p++) {
                        float4 c4;
                        if (direct) {
                            c4 = plb_load_clamped(pl.tex, gxh + p, gy);
                        } else {
                            float posx, posy;
                            plb_coord_pos(cc, gxh + p, gy, posx, posy);
                            c4 = pl.linear ? plb_tex_linear(pl.tex, posx, posy)
                                           : plb_tex_nearest(pl.tex, posx, posy);
                        }
                        const float tmp[4] = { c4.x * sc, c4.y * sc, c4.z * sc, c4.w * sc };
                        #pragma unroll
                        for (int k = 0; k < 4; k++) {
                            #pragma unroll
                            for (int j = 0; j < 4; j++)
                                col[p][j] = map[k] == j ? tmp[k] : col[p][j];
                        }
                    }
                }
                plb_run_light_ops_n<4>(t.prog, col);
                float4 px4[4];
                #pragma unroll
                for (int p = 0; p < 4; p++)
                    px4[p] = make_float4(col[p][0], col[p][1], col[p][2], col[p][3]);
                if (gxh + 4 <= t.out_w) {
                    plb_store4(t.dst, t.dst_x0 + gxh, t.dst_y0 + gy, px4);
                } else {
                    #pragma unroll
                    for (int p = 0; p < 4; p++) {
                        if (gxh + p < t.out_w)
                            plb_store(t.dst, t.dst_x0 + gxh + p, t.dst_y0 + gy, px4[p]);
                    }
                }
            }
        }
        return;
    }
    if constexpr (MODE == 2) {
        // Canonical merged plane pass of a YUV 4:2:0 HDR frame (host-checked, launch_up2):
        // chroma = the two EWA results -> colour[1], colour[2]; one r16 luma plane (direct,
        // debanded or AV1-grained) -> colour[0]; program = MATRIX (pl_shader_decode_color)
        // [+ peak-detection tap on PQ]; rgba16hf target. Straight-line, launch description
        // straight from the kernel parameters.
        const plb_tail_args &t = a.tail;
        float (*stage)[TH][TW] = (float (*)[TH][TW]) aux;
        #pragma unroll
        for (int oyo = 0; oyo < UP2_OY; oyo++) {
            #pragma unroll
            for (int k = 0; k < NC; k++) {
                #pragma unroll
                for (int half = 0; half < UP2_OX / 4; half++) {
                    float4 v;
                    v.x = a.norm[oyo][0] * acc[oyo][4 * half + 0][k];
                    v.y = a.norm[oyo][1] * acc[oyo][4 * half + 1][k];
                    v.z = a.norm[oyo][0] * acc[oyo][4 * half + 2][k];
                    v.w = a.norm[oyo][1] * acc[oyo][4 * half + 3][k];
                    *(float4 *) &stage[k][ty * UP2_OY + oyo][tx * UP2_OX + 4 * half] = v;
                }
            }
        }
        const bool has_peak = t.prog.n > 1;
        const plb_op_peak &pk = t.prog.ops[1].u.peak;
        const plb_color_op &mat = t.prog.ops[0];
        __shared__ plb_peak_smem ps;
        if (has_peak)
            plb_peak_begin(ps); // ends with a barrier
        else
            __syncthreads();
        const int wgs_x = has_peak ? TW / t.wg_w : 1, wgs_y = has_peak ? TH / t.wg_h : 1;
        const int nwgx = has_peak ? (t.out_w + t.wg_w - 1) / t.wg_w : 1;
        const int nwgy = has_peak ? (t.out_h + t.wg_h - 1) / t.wg_h : 1;
        // the tile is a whole number of work groups (host check), so their sizes are powers of two
        const int wg_sx = __ffs(t.wg_w) - 1, wg_sy = __ffs(t.wg_h) - 1;
        const int wg0x = X0 >> wg_sx, wg0y = Y0 >> wg_sy;
        const plb_plane_src &pl = t.planes[0];
        #pragma unroll 1
        for (int j = 0; j < UP2_OX * UP2_OY; j++) {
            const int idx = j * (UP2_TX * UP2_TY) + tid;
            const int lx = idx % TW, ly = idx / TW;
            const int gx = X0 + lx, gy = Y0 + ly;
            const bool inside = gx < t.out_w && gy < t.out_h;
            if (!has_peak && !inside)
                continue;
            float tmp[4];
            if (pl.kind == PLB_PLANE_GRAIN_AV1)
                plb_sample_plane_simple_t<0>(pl, t.plane_cc[0], gx, gy, t.frame_index, tmp);
            else
                plb_sample_plane_simple_t<1>(pl, t.plane_cc[0], gx, gy, t.frame_index, tmp);
            float color[4] = { tmp[0] * pl.scale, stage[0][ly][lx], stage[1][ly][lx], t.init[3] };
            plb_matv(mat.u.matrix.m, color[0], color[1], color[2]);
            if (mat.u.matrix.has_c) {
                color[0] += mat.u.matrix.c[0]; color[1] += mat.u.matrix.c[1]; color[2] += mat.u.matrix.c[2];
            }
            if (has_peak) {
                const uint32_t ypq = plb_peak_ypq<true>(TRC_PQ, pk, color);
                const int wlx = lx >> wg_sx, wly = ly >> wg_sy;
                const bool valid = (wg0x + wlx) < nwgx && (wg0y + wly) < nwgy;
                plb_peak_add(ps, pk, valid, wly * wgs_x + wlx, ypq);
            }
            if (inside) {
                const __half2 lo = __floats2half2_rn(color[0], color[1]), hi = __floats2half2_rn(color[2], color[3]);
                uint2 raw;
                raw.x = *(const uint32_t *) &lo;
                raw.y = *(const uint32_t *) &hi;
                *(uint2 *) ((char *) t.dst.ptr + (size_t) (t.dst_y0 + gy) * t.dst.pitch +
                            (size_t) (t.dst_x0 + gx) * 8) = raw;
            }
        }
        if (has_peak)
            plb_peak_finish(ps, pk, wgs_x, wgs_y, X0 / t.wg_w, Y0 / t.wg_h, nwgx, nwgy,
                            (uint32_t) (t.wg_w * t.wg_h));
        return;
    } else {

    const plb_tail_args &t = plb_s_tail;
    bool has_peak = false;
    for (int i = 0; i < plb_s_prog.n; i++)
        has_peak |= plb_s_prog.ops[i].kind == PLB_OP_PEAK;
    const int wgs_x = has_peak ? TW / t.wg_w : 1, wgs_y = has_peak ? TH / t.wg_h : 1;
    const int nwgx = has_peak ? (t.out_w + t.wg_w - 1) / t.wg_w : 1;
    const int nwgy = has_peak ? (t.out_h + t.wg_h - 1) / t.wg_h : 1;
    __shared__ plb_peak_smem ps;
    if (has_peak)
        plb_peak_begin(ps);
    const plb_op_peak *pk_any = nullptr;
    #pragma unroll
    for (int oyo = 0; oyo < UP2_OY; oyo++) {
        const int gy = Y0 + ty * UP2_OY + oyo;
        #pragma unroll
        for (int oxo = 0; oxo < UP2_OX; oxo++) {
            const int gx = X0 + tx * UP2_OX + oxo;
            const bool inside = gx < t.out_w && gy < t.out_h;
            if (!has_peak && !inside)
                continue;
            const float norm = a.norm[oyo][oxo & 1];
            float prim[4] = {0.0f, 0.0f, 0.0f, 1.0f};
            #pragma unroll
            for (int k = 0; k < NC; k++)
                prim[k] = norm * acc[oyo][oxo][k];
            float color[4];
            uint32_t ypq = 0;
            const plb_op_peak *pk = nullptr;
            plb_tail_color(gx, gy, prim, color, &ypq, &pk);
            if (has_peak) {
                const int lx = (gx - X0) / t.wg_w, ly = (gy - Y0) / t.wg_h;
                const bool valid = (X0 / t.wg_w + lx) < nwgx && (Y0 / t.wg_h + ly) < nwgy;
                pk_any = pk;
                plb_peak_add(ps, *pk, valid, ly * wgs_x + lx, ypq);
            }
            if (inside && t.dst.ptr)
                plb_store(t.dst, t.dst_x0 + gx, t.dst_y0 + gy,
                          make_float4(color[0], color[1], color[2], color[3]));
        }
    }
    if (has_peak)
        plb_peak_finish(ps, *pk_any, wgs_x, wgs_y, X0 / t.wg_w, Y0 / t.wg_h, nwgx, nwgy,
                        (uint32_t) (t.wg_w * t.wg_h));
    } // MODE != 2
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

// Tap list with the reference's build-time culling (sampling.c:508-522, :776-783)
static int build_taps(float radius, float radius_zero, bool use_ar, int8_t *tx, int8_t *ty,
                      uint8_t *flags)
{
    const int bound = (int) ceilf(radius);
    int n = 0;
    for (int y = 1 - bound; y <= bound; y++) {
        for (int x = 1 - bound; x <= bound; x++) {
            int yy = y > 0 ? y - 1 : y;
            int xx = x > 0 ? x - 1 : x;
            float dmin = sqrt((double) (xx * xx + yy * yy));
            if (dmin >= radius)
                continue;
            if (n >= MAX_TAPS)
                return -1;
            tx[n] = x;
            ty[n] = y;
            flags[n] = (dmin >= radius - M_SQRT2 ? 1 : 0) |
                       (use_ar && dmin < radius_zero ? 2 : 0);
            n++;
        }
    }
    return n;
}

static float host_lut1d(const float *lut, int n, float x)
{
    float fpos = fminf(fmaxf(x, 0.0f), 1.0f) * (float) (n - 1);
    float fbase = floorf(fpos), fceil = ceilf(fpos);
    float t = fpos - fbase;
    return lut[(int) fbase] * (1.0f - t) + lut[(int) fceil] * t;
}

#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)

// cuTensorMapEncodeTiled through the runtime's driver entry point table (no link against libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_tiled()
{
    static encode_tiled_fn fn;
    static int tried;
    if (!tried) {
        void *p = NULL;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn) p;
        tried = 1;
    }
    return fn;
}

// 2-D tensor map over a pitched texture of `texel`-byte texels (8: rgba16hf, 4: rg16hf), box =
// the CTA's source tile. false: TMA unavailable for this texture (alignment), use the loader.
static bool make_tile_map(CUtensorMap *map, const plb_tex_desc &src, int texel)
{
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc || ((uintptr_t) src.ptr & 15) || (src.pitch & 15) || src.w < UP2_TILE_W || src.h < UP2_TILE_H)
        return false;
    const cuuint64_t dims[2] = { (cuuint64_t) src.w, (cuuint64_t) src.h };
    const cuuint64_t strides[1] = { (cuuint64_t) src.pitch };
    const cuuint32_t box[2] = { UP2_TILE_W, UP2_TILE_H };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUtensorMapDataType dt = texel == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    return enc(map, dt, 2, src.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NC, int XM, int YM, int MODE>
static void launch_up2_m(pl_gpu gpu, dim3 grid, dim3 block, polar_up2_args &args)
{
    // dedicated loaders (and TMA for interior tiles) for the two FBO layouts of the hot path
    const plb_tex_desc &s = args.src;
    args.src_kind = s.ptr2 || s.addr ? UP2_SRC_ANY
                  : (s.kind == PLB_FK_HALF && s.comps == 4) ? UP2_SRC_HALF4
                  : (s.kind == PLB_FK_HALF && s.comps == 2 && NC <= 2) ? UP2_SRC_HALF2 : UP2_SRC_ANY;
    static int tma_on = -1; // PLB_UP2_TMA=0: manual loader everywhere (A/B measurements)
    if (tma_on < 0) {
        const char *e = getenv("PLB_UP2_TMA");
        tma_on = e ? atoi(e) != 0 : 1;
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    args.use_tma = args.src_kind != UP2_SRC_ANY && tma_on &&
                   make_tile_map(&map, s, args.src_kind == UP2_SRC_HALF4 ? 8 : 4);
    constexpr int smem = up2_smem_bytes<NC, MODE>();
    static bool attr_set; // per instantiation; the attribute is per function and device-wide
    if (!attr_set) {
        cudaFuncSetAttribute(k_polar_up2<NC, XM, YM, MODE>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    k_polar_up2<NC, XM, YM, MODE><<<grid, block, smem, STREAM(gpu)>>>(args, map);
}

template <int NC, int XM, int YM>
static bool launch_up2(pl_gpu gpu, polar_up2_args &args)
{
    dim3 block(UP2_TX, UP2_TY);
    // the tile is a whole number of peak work groups, so this also covers overhanging ones
    dim3 grid((args.tail.out_w + UP2_TW - 1) / UP2_TW, (args.tail.out_h + UP2_TH - 1) / UP2_TH);
    // light pass: only directly sampled planes to merge and at most a "light" colour program
    bool light = true;
    for (int i = 0; i < args.tail.prog.n; i++) {
        const int k = args.tail.prog.ops[i].kind;
        light &= k == PLB_OP_SCALE || k == PLB_OP_MATRIX || k == PLB_OP_ROUND_FBO ||
                 k == PLB_OP_ALPHA || k == PLB_OP_SWIZZLE;
    }
    for (int i = 0; i < args.tail.nplanes; i++)
        light &= args.tail.planes[i].kind == PLB_PLANE_DIRECT;
    const bool pure = light && args.tail.dst.ptr;
    // canonical merged YUV 4:2:0 plane pass (MODE 2, see the kernel): EWA chroma -> colour[1..2],
    // one r16 luma plane (direct 1:1 / debanded / AV1-grained) -> colour[0], program = MATRIX
    // [PEAK on PQ], rgba16hf target; not in exact-math mode (static curves)
    const plb_tail_args &tl = args.tail;
    bool hdr = NC == 2 && !pure && tl.dst.ptr && tl.dst.kind == PLB_FK_HALF && tl.dst.comps == 4 &&
               !plb_gpu_priv(gpu)->exact_math && tl.nplanes == 1 &&
               tl.src_map[0] == 1 && tl.src_map[1] == 2 &&
               tl.planes[0].map[0] == 0 && tl.planes[0].map[1] < 0 && tl.planes[0].map[2] < 0 &&
               tl.planes[0].map[3] < 0 && (tl.prog.n == 1 || tl.prog.n == 2) &&
               tl.prog.ops[0].kind == PLB_OP_MATRIX &&
               (tl.prog.n == 1 || (tl.prog.ops[1].kind == PLB_OP_PEAK && tl.prog.ops[1].u.peak.lin.trc == TRC_PQ));
    if (hdr) {
        const plb_plane_src &pl = tl.planes[0];
        const bool r16 = pl.tex.kind == PLB_FK_UNORM16 && pl.tex.comps == 1 && !pl.tex.ptr2 && !pl.tex.addr;
        hdr = (pl.kind == PLB_PLANE_DIRECT && r16 && tl.plane_cc[0].ident) ||
              (pl.kind == PLB_PLANE_DEBAND && r16) || pl.kind == PLB_PLANE_GRAIN_AV1;
    }
    static int hdr_on = -1; // PLB_UP2_HDR=0 keeps the general tail (A/B measurements)
    if (hdr_on < 0) {
        const char *e = getenv("PLB_UP2_HDR");
        hdr_on = e ? atoi(e) != 0 : 1;
    }
    if (pure)
        launch_up2_m<NC, XM, YM, 1>(gpu, grid, block, args);
    else if (hdr && hdr_on)
        launch_up2_m<NC, XM, YM, NC == 2 ? 2 : 0>(gpu, grid, block, args);
    else
        launch_up2_m<NC, XM, YM, 0>(gpu, grid, block, args);
    return plb_cuda_check(gpu, cudaGetLastError(), "k_polar_up2");
}

extern "C" bool plb_launch_polar(pl_gpu gpu, const struct plb_polar_launch *p)
{
    struct plb_gpu *g = plb_gpu_priv(gpu);
    const struct plb_pixel_launch *tl = &p->tail;
    const int bound = (int) ceilf(p->radius);
    if (bound > 6) {
        PLB_ERR(gpu->log, "polar radius %f exceeds implementation capacity", p->radius);
        return false;
    }
    bool has_peak = false;
    for (int i = 0; i < tl->prog.n; i++)
        has_peak |= tl->prog.ops[i].kind == PLB_OP_PEAK;
    const int wg_w = PLB_DEF(tl->wg_w, 32), wg_h = PLB_DEF(tl->wg_h, 32);

    // --- 2x fast path eligibility ---
    // whole texture, origin 0 (centred phases) or 0.25 (left/top sited chroma) per axis
    const int xm = p->rect.x0 == 0.0f ? 0 : p->rect.x0 == 0.25f ? 1 : -1;
    const int ym = p->rect.y0 == 0.0f ? 0 : p->rect.y0 == 0.25f ? 1 : -1;
    const bool full = xm >= 0 && ym >= 0 && p->rect.x1 - p->rect.x0 == (float) p->src.w &&
                      p->rect.y1 - p->rect.y0 == (float) p->src.h;
    const uint8_t mask = p->comp_mask;
    const int nc = mask == 0x1 ? 1 : mask == 0x3 ? 2 : mask == 0x7 ? 3 : mask == 0xF ? 4 : 0;
    bool fast = p->allow_fast && tl->coord_mode == PLB_COORD_EXACT && full && p->host_lut &&
                tl->out_w == 2 * p->src.w && tl->out_h == 2 * p->src.h && bound <= 4 &&
                !(p->antiring > 0) && nc && nc <= p->src.comps &&
                (nc <= 2 || (xm == 0 && ym == 0));
    const int TW = UP2_TW, TH = UP2_TH;
    if (has_peak && (TW % wg_w || TH % wg_h || (TW / wg_w) * (TH / wg_h) > PLB_MAX_WG_PER_CTA))
        fast = false;
    if (has_peak && !fast && (32 % wg_w || 32 % wg_h)) {
        PLB_ERR(gpu->log, "peak detection: unsupported work group size %dx%d", wg_w, wg_h);
        return false;
    }

    if (fast) {
        polar_up2_args args;
        memset(&args, 0, sizeof(args));
        args.src = p->src;
        plb_fill_tail(&args.tail, tl);
        args.tail.wg_w = wg_w;
        args.tail.wg_h = wg_h;
        memcpy(args.tail.src_map, p->src_map, sizeof(p->src_map));
        int8_t tx[MAX_TAPS], ty[MAX_TAPS];
        uint8_t fl[MAX_TAPS];
        const int ntaps = build_taps(p->radius, p->radius_zero, false, tx, ty, fl);
        const float inv_radius = 1.0f / p->radius;
        int max_d16 = 0;
        for (int py = 0; py < 2; py++) {
            for (int px = 0; px < 2; px++) {
                const float fx = up2_frac4(xm, px) * 0.25f, fy = up2_frac4(ym, py) * 0.25f;
                float wsum = 0.0f;
                for (int t = 0; t < ntaps; t++) {
                    const float dx = (float) tx[t] - fx, dy = (float) ty[t] - fy;
                    const float d = sqrtf(fmaf(dx, dx, dy * dy));
                    if ((fl[t] & 1) && !(d < p->radius))
                        continue;
                    const float w = host_lut1d(p->host_lut, p->lut_entries, d * inv_radius);
                    wsum += w;
                    args.w[py][px][ty[t] + 3][tx[t] + 3] = w;
                    if (w != 0.0f) {
                        int ax = 4 * tx[t] - up2_frac4(xm, px), ay = 4 * ty[t] - up2_frac4(ym, py);
                        int d16 = ax * ax + ay * ay;
                        max_d16 = PLB_MAX(max_d16, d16);
                    }
                }
                args.norm[py][px] = p->scale / wsum;
            }
        }
        // the phase kernels cull taps beyond radius^2 = 10 at compile time (EWA-Lanczos and its
        // sharpened variants: 30 of 64 taps remain); wider filters take the generic kernel
        if (max_d16 < 160) {
            bool ok;
            // centred phases for every component count; sited-chroma phases for 1-2 components
            const int mode = xm * 2 + ym;
            if (mode == 0) {
                switch (nc) {
                case 1: ok = launch_up2<1, 0, 0>(gpu, args); break;
                case 2: ok = launch_up2<2, 0, 0>(gpu, args); break;
                case 3: ok = launch_up2<3, 0, 0>(gpu, args); break;
                default: ok = launch_up2<4, 0, 0>(gpu, args); break;
                }
            } else if (nc == 1) {
                ok = mode == 1 ? launch_up2<1, 0, 1>(gpu, args)
                   : mode == 2 ? launch_up2<1, 1, 0>(gpu, args)
                               : launch_up2<1, 1, 1>(gpu, args);
            } else {
                ok = mode == 1 ? launch_up2<2, 0, 1>(gpu, args)
                   : mode == 2 ? launch_up2<2, 1, 0>(gpu, args)
                               : launch_up2<2, 1, 1>(gpu, args);
            }
            g->launches++;
            return ok;
        }
        if (has_peak && (32 % wg_w || 32 % wg_h)) {
            PLB_ERR(gpu->log, "peak detection: unsupported work group size %dx%d", wg_w, wg_h);
            return false;
        }
    }

    polar_generic_args a;
    memset(&a, 0, sizeof(a));
    a.src = p->src;
    a.cc = plb_coord_init(tl->coord_mode, p->src, p->rect, tl->out_w, tl->out_h);
    a.lut = p->lut;
    a.lut_entries = p->lut_entries;
    a.radius = p->radius;
    a.inv_radius = 1.0f / p->radius;
    a.radius_zero = p->radius_zero;
    a.antiring = p->antiring;
    a.scale = p->scale;
    a.comp_mask = p->comp_mask;
    a.ntaps = build_taps(p->radius, p->radius_zero, p->antiring > 0, a.tap_x, a.tap_y, a.tap_flags);
    if (a.ntaps < 0) {
        PLB_ERR(gpu->log, "polar filter needs too many taps");
        return false;
    }
    plb_fill_tail(&a.tail, tl);
    a.tail.wg_w = wg_w;
    a.tail.wg_h = wg_h;
    memcpy(a.tail.src_map, p->src_map, sizeof(p->src_map));
    dim3 block(32, 8);
    dim3 grid((tl->out_w + 31) / 32, (tl->out_h + 31) / 32);
    if (has_peak) {
        grid.x = ((tl->out_w + wg_w - 1) / wg_w * wg_w + 31) / 32;
        grid.y = ((tl->out_h + wg_h - 1) / wg_h * wg_h + 31) / 32;
    }
    k_polar_generic<<<grid, block, 0, STREAM(gpu)>>>(a);
    g->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_polar_generic");
}

// ---------------------------------------------------------------------------------------
// pl_tex_clear
// ---------------------------------------------------------------------------------------

__global__ void k_clear(plb_tex_desc dst, float4 color)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < dst.w && y < dst.h)
        plb_store(dst, x, y, color);
}

extern "C" bool plb_launch_clear(pl_gpu gpu, const struct plb_tex_desc *dst, const float color[4])
{
    dim3 block(32, 8), grid((dst->w + 31) / 32, (dst->h + 7) / 8);
    k_clear<<<grid, block, 0, STREAM(gpu)>>>(*dst, make_float4(color[0], color[1], color[2], color[3]));
    plb_gpu_priv(gpu)->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_clear");
}
