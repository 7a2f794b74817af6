// EWA 2x phase kernels (k_polar_up2) and their launcher, shared by the translation units that
// instantiate them (k_up2_*.cu: split by component count so that nvcc compiles them in parallel).
// See k_polar.cu for the generic polar kernel and the dispatch (plb_launch_polar).
#pragma once
#include "k_up2_args.cuh"

#include <cuda.h> // CUtensorMap (driver API types; the entry point is resolved at run time)

// ---------------------------------------------------------------------------------------
// 2x fast path
// ---------------------------------------------------------------------------------------

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
enum { UP2_SRC_ANY = 0, UP2_SRC_HALF4 = 1, UP2_SRC_HALF2 = 2, UP2_SRC_R16X2 = 3 /* planar pair of r16 */ };

// the generic fetch is large (run-time format / address-mode dispatch): out of line, so that the
// tile loader of the hot layouts stays small
static __device__ __noinline__ float4 up2_load_any(const plb_tex_desc &t, int x, int y)
{
    return plb_load_clamped(t, x, y);
}

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
// Thread = 8 x 2 outputs (four source cells) from a 12 x 9 source window. Components are
// processed in PAIRS with the packed fp32 FMA of sm_100 (FFMA2: two IEEE fmas per instruction,
// the weight broadcast from a uniform register): NC = 2 -> one FFMA2 per tap, 3 -> FFMA2 + FFMA,
// 4 -> two FFMA2 — the kernel is bound by instruction fetch / issue, not by the FMA pipe, so
// halving the instruction stream is what counts; results are bit-identical to scalar fmas.
// The source tile is staged as float2 planes (component pairs) plus a float plane for an odd
// component. A float2 row is split into even and odd 16-byte chunks (texel pairs 0-1, 4-5, ...
// and 2-3, 6-7, ...): thread tx reads chunks 2 tx + q, so within each half consecutive lanes
// read consecutive 16 bytes (conflict-free LDS.128).
template <int NC, int XM, int YM, int MODE>
__global__ void __launch_bounds__(UP2_TX * UP2_TY, MODE == 1 ? 3 : 2)
k_polar_up2(const __grid_constant__ polar_up2_args a, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int R2_16 = 160; // taps beyond radius^2 = 10 never contribute (host check)
    constexpr bool PURE = MODE == 1;
    constexpr int TW = UP2_TW, TH = UP2_TH;
    constexpr int TILE_BYTES = NC * UP2_TILE_H * UP2_TILE_W * 4;
    constexpr int NP2 = NC / 2, NP1 = NC & 1;          // float2 planes, float planes
    constexpr int ROW2 = UP2_TILE_W * 8, HALF2 = ROW2 / 2; // bytes of a float2 row / of its even half
    const int SRC = a.src_kind; // launch-uniform
    // the raw TMA tile is dead once converted; the MODE 2 staging buffer reuses its space
    extern __shared__ __align__(128) unsigned char smem[]; // up2_smem_bytes<NC, MODE>()
    __shared__ __align__(8) uint64_t bar;
    unsigned char *tile2 = smem;                                    // [NP2][TILE_H] rows of ROW2 bytes
    float (*tile1)[UP2_TILE_W] = (float (*)[UP2_TILE_W]) (smem + NP2 * UP2_TILE_H * ROW2); // [TILE_H][TILE_W]
    unsigned char *aux = smem + TILE_BYTES;
    // texel (r, c) of the tile <- the NC components of v
    auto put = [&](int r, int c, float4 v) {
        const int j = c >> 1;
        unsigned char *row = tile2 + (size_t) r * ROW2 + ((j & 1) ? HALF2 : 0) + (j >> 1) * 16 + (c & 1) * 8;
        if (NP2 >= 1)
            *(float2 *) row = make_float2(v.x, v.y);
        if (NP2 >= 2)
            *(float2 *) (row + (size_t) UP2_TILE_H * ROW2) = make_float2(v.z, v.w);
        if (NP1)
            tile1[r][c] = NC == 1 ? v.x : v.z;
    };

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

    const bool interior = a.use_tma && col0 >= 0 && row0 >= 0 &&
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
                put(r, c, make_float4(lo.x, lo.y, hi.x, hi.y));
            } else {
                const float2 v = __half22float2(((const __half2 *) aux)[i]);
                put(r, c, make_float4(v.x, v.y, 0.0f, 1.0f));
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
            } else if (SRC == UP2_SRC_R16X2) {
                const int x = plb_clampi(col0 + c, 0, a.src.w - 1), y = plb_clampi(row0 + r, 0, a.src.h - 1);
                const unsigned u0 = *(const uint16_t *) ((const char *) a.src.ptr + (size_t) y * a.src.pitch + 2 * x);
                const unsigned u1 = *(const uint16_t *) ((const char *) a.src.ptr2 + (size_t) y * a.src.pitch2 + 2 * x);
                v = make_float4(plb_unorm16_to_float(u0), plb_unorm16_to_float(u1), 0.0f, 1.0f);
            } else {
                v = up2_load_any(a.src, col0 + c, row0 + r);
            }
            put(r, c, v);
        }
    }
    __syncthreads();

    float2 acc2[UP2_OY][UP2_OX][NP2 ? NP2 : 1];
    float acc1[UP2_OY][UP2_OX];
    #pragma unroll
    for (int i = 0; i < UP2_OY; i++)
        #pragma unroll
        for (int j = 0; j < UP2_OX; j++) {
            #pragma unroll
            for (int k = 0; k < (NP2 ? NP2 : 1); k++)
                acc2[i][j][k] = make_float2(0.0f, 0.0f);
            acc1[i][j] = 0.0f;
        }
    // component k of output (i, j)
#define ACC(i, j, k) ((k) < 2 * NP2 ? (((k) & 1) ? acc2[i][j][(k) >> 1].y : acc2[i][j][(k) >> 1].x) : acc1[i][j])

    // thread's source cells: columns i0 = X0/2 + 4*tx (+0..3), row jj = Y0/2 + ty
    // window column wx <-> source column i0 - 4 + wx  (tile column 4*tx + wx)
    // window row    wy <-> source row    jj - 4 + wy  (tile row    ty + wy)
    #pragma unroll
    for (int wy = 0; wy < 9; wy++) {
        float2 v2[NP2 ? NP2 : 1][12];
        float v1[12];
        #pragma unroll
        for (int p = 0; p < NP2; p++) {
            const unsigned char *row = tile2 + (size_t) (p * UP2_TILE_H + ty + wy) * ROW2 + tx * 16;
            #pragma unroll
            for (int q = 0; q < 3; q++) {
                const float4 e = *(const float4 *) (row + q * 16);         // texels 4 q, 4 q + 1
                const float4 o = *(const float4 *) (row + HALF2 + q * 16); // texels 4 q + 2, 4 q + 3
                v2[p][4 * q] = make_float2(e.x, e.y);
                v2[p][4 * q + 1] = make_float2(e.z, e.w);
                v2[p][4 * q + 2] = make_float2(o.x, o.y);
                v2[p][4 * q + 3] = make_float2(o.z, o.w);
            }
        }
        if (NP1) {
            const float4 *rowp = (const float4 *) &tile1[ty + wy][4 * tx];
            #pragma unroll
            for (int q = 0; q < 3; q++) {
                const float4 f = rowp[q];
                v1[4 * q] = f.x;
                v1[4 * q + 1] = f.y;
                v1[4 * q + 2] = f.z;
                v1[4 * q + 3] = f.w;
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
                        for (int p = 0; p < NP2; p++)
                            acc2[oyo][oxo][p] = __ffma2_rn(make_float2(w, w), v2[p][wx], acc2[oyo][oxo][p]);
                        if (NP1)
                            acc1[oyo][oxo] = __fmaf_rn(w, v1[wx], acc1[oyo][oxo]);
                    }
                }
            }
        }
    }

    if (PURE) {
        // Canonical light pass (host-checked, launch_up2): the NC results go to colour[base ..
        // base + NC), optionally one directly sampled r16 luma plane to colour[0]; the colour
        // program is [matrix] [scale] [identity swizzle], folded into flags; vector stores.
        const plb_tail_args &t = a.tail;
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
                const int gxh = gx + 4 * half;
                float col[4][4];
                float luma[4] = { t.init[0], t.init[0], t.init[0], t.init[0] };
                if (a.ep_luma) {
                    // four consecutive texels of the r16 plane (gxh % 4 == 0: one 8-byte load)
                    const plb_plane_src &pl = t.planes[0];
                    const uint16_t *row = (const uint16_t *) ((const char *) pl.tex.ptr + (size_t) gy * pl.tex.pitch);
                    if (gxh + 4 <= pl.tex.w) {
                        const uint2 raw = *(const uint2 *) (row + gxh);
                        luma[0] = plb_unorm16_to_float(raw.x & 0xFFFFu) * pl.scale;
                        luma[1] = plb_unorm16_to_float(raw.x >> 16) * pl.scale;
                        luma[2] = plb_unorm16_to_float(raw.y & 0xFFFFu) * pl.scale;
                        luma[3] = plb_unorm16_to_float(raw.y >> 16) * pl.scale;
                    } else {
                        #pragma unroll
                        for (int p = 0; p < 4; p++)
                            luma[p] = plb_unorm16_to_float(row[min(gxh + p, pl.tex.w - 1)]) * pl.scale;
                    }
                }
                #pragma unroll
                for (int p = 0; p < 4; p++) {
                    const float norm = a.norm[oyo][p & 1];
                    #pragma unroll
                    for (int j = 0; j < 4; j++)
                        col[p][j] = t.init[j];
                    if (a.ep_base) { // chroma pair next to a luma plane
                        col[p][0] = luma[p];
                        #pragma unroll
                        for (int k = 0; k < NC && k < 3; k++)
                            col[p][1 + k] = norm * ACC(oyo, 4 * half + p, k);
                    } else {
                        #pragma unroll
                        for (int k = 0; k < NC; k++)
                            col[p][k] = norm * ACC(oyo, 4 * half + p, k);
                        #pragma unroll
                        for (int k = NC; k < 4; k++) { // unsampled components mapped onto the colour
                            if (a.ep_fill & (1 << k))
                                col[p][k] = k == 3 ? 1.0f : 0.0f;
                        }
                    }
                    if (a.ep_has_mat) {
                        plb_matv(a.ep_mat, col[p][0], col[p][1], col[p][2]);
                        if (a.ep_mat_has_c) {
                            col[p][0] += a.ep_mat_c[0]; col[p][1] += a.ep_mat_c[1]; col[p][2] += a.ep_mat_c[2];
                        }
                    }
                    if (a.ep_has_scale) {
                        col[p][0] *= a.ep_scale; col[p][1] *= a.ep_scale; col[p][2] *= a.ep_scale;
                        if (a.ep_scale_alpha)
                            col[p][3] *= a.ep_scale;
                    }
                    if (a.ep_has_swz) {
                        const float o3 = col[p][3];
                        if (!(a.ep_swz_keep & 1)) col[p][0] = 0.0f;
                        if (!(a.ep_swz_keep & 2)) col[p][1] = 0.0f;
                        if (!(a.ep_swz_keep & 4)) col[p][2] = 0.0f;
                        if (!(a.ep_swz_keep & 8)) col[p][3] = 1.0f;
                        if (a.ep_swz_force_alpha)
                            col[p][3] = o3;
                    }
                }
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
                    v.x = a.norm[oyo][0] * ACC(oyo, 4 * half + 0, k);
                    v.y = a.norm[oyo][1] * ACC(oyo, 4 * half + 1, k);
                    v.z = a.norm[oyo][0] * ACC(oyo, 4 * half + 2, k);
                    v.w = a.norm[oyo][1] * ACC(oyo, 4 * half + 3, k);
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
                prim[k] = norm * ACC(oyo, oxo, k);
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
    args.src_kind = s.addr ? UP2_SRC_ANY
                  : (s.ptr2 && s.kind == PLB_FK_UNORM16 && NC == 2) ? UP2_SRC_R16X2
                  : s.ptr2 ? UP2_SRC_ANY
                  : (s.kind == PLB_FK_HALF && s.comps == 4) ? UP2_SRC_HALF4
                  : (s.kind == PLB_FK_HALF && s.comps == 2 && NC <= 2) ? UP2_SRC_HALF2 : UP2_SRC_ANY;
    static int tma_on = -1; // PLB_UP2_TMA=0: manual loader everywhere (A/B measurements)
    if (tma_on < 0) {
        const char *e = getenv("PLB_UP2_TMA");
        tma_on = e ? atoi(e) != 0 : 1;
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    args.use_tma = (args.src_kind == UP2_SRC_HALF4 || args.src_kind == UP2_SRC_HALF2) && tma_on &&
                   make_tile_map(&map, s, args.src_kind == UP2_SRC_HALF4 ? 8 : 4);
    constexpr int smem = up2_smem_bytes<NC, MODE>();
    // per instantiation and per device (the attribute belongs to the function in one context)
    static unsigned long long attr_devs;
    const int dev = plb_gpu_priv(gpu)->device & 63;
    if (!((attr_devs >> dev) & 1)) {
        cudaFuncSetAttribute(k_polar_up2<NC, XM, YM, MODE>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_devs |= 1ull << dev;
    }
    k_polar_up2<NC, XM, YM, MODE><<<grid, block, smem, STREAM(gpu)>>>(args, map);
}

template <int NC, int XM, int YM>
static bool launch_up2(pl_gpu gpu, polar_up2_args &args)
{
    dim3 block(UP2_TX, UP2_TY);
    // the tile is a whole number of peak work groups, so this also covers overhanging ones
    dim3 grid((args.tail.out_w + UP2_TW - 1) / UP2_TW, (args.tail.out_h + UP2_TH - 1) / UP2_TH);
    // canonical light pass (MODE 1, see the kernel): results -> colour[base .. base + NC), at most
    // one directly sampled r16 luma plane -> colour[0], program = [matrix] [scale] [identity swizzle]
    bool pure = args.tail.dst.ptr != NULL;
    {
        const plb_tail_args &tl = args.tail;
        // component k of the sampler goes to colour[src_map[k]]; components beyond the sampled
        // ones carry the defaults (0, 0, 0, 1). Accepted: base + k for the sampled ones, and for the
        // rest "unused" or (base 0 only) the identity, which stores the default into the colour.
        int base = -1;
        for (int b0 = 0; b0 < 2 && base < 0; b0++) {
            bool m = b0 + NC <= 4;
            int fill = 0;
            for (int k = 0; k < 4; k++) {
                if (k < NC) {
                    m &= tl.src_map[k] == b0 + k;
                } else if (tl.src_map[k] == k && b0 == 0) {
                    fill |= 1 << k;
                } else {
                    m &= tl.src_map[k] < 0;
                }
            }
            if (m) {
                base = b0;
                args.ep_fill = fill;
            }
        }
        pure &= base >= 0 && tl.nplanes <= 1 && (base == 0 ? tl.nplanes == 0 : true);
        if (pure && tl.nplanes == 1) {
            const plb_plane_src &pl = tl.planes[0];
            pure = base == 1 && pl.kind == PLB_PLANE_DIRECT && tl.plane_cc[0].ident &&
                   pl.tex.kind == PLB_FK_UNORM16 && pl.tex.comps == 1 && !pl.tex.ptr2 && !pl.tex.addr &&
                   !((uintptr_t) pl.tex.ptr & 7) && !(pl.tex.pitch & 7) &&
                   pl.map[0] == 0 && pl.map[1] < 0 && pl.map[2] < 0 && pl.map[3] < 0;
            args.ep_luma = pure;
        } else if (pure && base == 1) {
            pure = false; // chroma without its luma plane: general tail
        }
        args.ep_base = base == 1;
        int i = 0;
        if (pure && i < tl.prog.n && tl.prog.ops[i].kind == PLB_OP_MATRIX) {
            args.ep_has_mat = 1;
            memcpy(args.ep_mat, tl.prog.ops[i].u.matrix.m, sizeof(args.ep_mat));
            memcpy(args.ep_mat_c, tl.prog.ops[i].u.matrix.c, sizeof(args.ep_mat_c));
            args.ep_mat_has_c = tl.prog.ops[i].u.matrix.has_c;
            i++;
        }
        if (pure && i < tl.prog.n && tl.prog.ops[i].kind == PLB_OP_SCALE) {
            args.ep_has_scale = 1;
            args.ep_scale = tl.prog.ops[i].u.scale.s;
            args.ep_scale_alpha = tl.prog.ops[i].u.scale.with_alpha;
            i++;
        }
        if (pure && i < tl.prog.n && tl.prog.ops[i].kind == PLB_OP_SWIZZLE) {
            const plb_color_op &op = tl.prog.ops[i];
            args.ep_has_swz = 1;
            args.ep_swz_force_alpha = op.u.swizzle.force_alpha;
            for (int k = 0; k < 4; k++) {
                const int m = op.u.swizzle.map[k];
                const int sel = (k < op.u.swizzle.comps && m >= 0) ? m : -1;
                if (sel == k)
                    args.ep_swz_keep |= 1 << k;
                else if (sel >= 0)
                    pure = false; // a real permutation
            }
            i++;
        }
        pure &= i == tl.prog.n;
    }
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

