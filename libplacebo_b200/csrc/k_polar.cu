// Polar (EWA) resampling — the CUDA replacement of the GLSL emitted by
// pl_shader_sample_polar (src/shaders/sampling.c:587-912, per-tap body polar_sample :503-558).
//
// Two kernels:
//
//  k_polar_generic<HALF4>  any ratio / rect / filter / anti-ringing. One thread per output pixel,
//                    per-pixel sub-texel fraction and per-tap weight (sqrt + division + LUT
//                    lerp), source texels straight from L1/L2 (HALF4: rgba16hf image without
//                    anti-ringing, format-specialised fetch). Follows the oracle operation by
//                    operation in both coordinate modes, so it is bit-exact with it up to store
//                    rounding.
//
//  k_polar_up2<NC,XM,YM,MODE> (k_up2.cuh, instantiated in k_up2_{a,b,c}.cu)
//                    exact 2x upscale of a whole texture (chroma 4:2:0 -> luma size, 1080p ->
//                    4K): with exact coordinates the sub-texel fraction takes only two values
//                    per axis ({0.75, 0.25} centred, {0, 0.5} for left / top sited chroma), so
//                    the 4 x 64 tap weights are tabulated once on the host (same fp32 ops as the
//                    per-pixel path -> identical values) and handed to the kernel as
//                    __grid_constant__ parameters. CTA = 16x16 threads, each producing an 8x2
//                    output block of a 128x32 tile; the 72x24 source tile is fetched by TMA
//                    (cp.async.bulk.tensor.2d + mbarrier; border tiles by clamped loads) and
//                    repacked into float2 pair-planes whose even / odd 16-byte chunks are split
//                    for conflict-free LDS.128; taps are visited in the reference's (y, x)
//                    order and accumulated with packed FFMA2, so results equal the
//                    exact-coordinate oracle bit for bit. Taps that can never contribute for
//                    the filter's radius are removed at compile time (30 of 64 for EWA-Lanczos).
//                    MODE 0: general tail (interpreter); 1: folded "scale + identity swizzle"
//                    epilogue; 2: the merged HDR plane pass (deband / grain / decode / PQ peak
//                    tap inlined, run x-fastest from a shared-memory staging of the results).
//
// Roofline notes: 2x EWA-Lanczos needs 30 taps x NC FMA per output pixel; at
// 128 FMA/clk/SM that is the binding limit (not HBM): 0.70 clk/px/SM for NC = 3.
#include "k_up2_args.cuh"

void plb_fill_tail(plb_tail_args *t, const struct plb_pixel_launch *p);

PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_polar)

#define MAX_TAPS 144

struct polar_generic_args {
    plb_tex_desc src;
    plb_coord cc;
    const float *lut;
    int lut_entries;
    float radius, radius_zero, antiring, scale;
    uint8_t comp_mask;
    int ntaps;
    int8_t tap_x[MAX_TAPS], tap_y[MAX_TAPS];
    uint8_t tap_flags[MAX_TAPS]; // bit0 guard, bit1 anti-ringing candidate
    plb_tail_args tail;
};

// The same sum for the usual image source — an rgba16hf texture with clamp addressing, no
// anti-ringing: format-specialised fetch, no per-component tests. Same operations in the same
// order as polar_pixel(), hence the same values.
PLB_DEV void polar_pixel_half4(const polar_generic_args &a, int gx, int gy, float out[4])
{
    int jx, jy;
    float fx, fy;
    plb_coord_texel(a.cc, gx, gy, jx, jy, fx, fy);
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f, wsum = 0.0f;
    const int wm1 = a.src.w - 1, hm1 = a.src.h - 1;
    #pragma unroll 2
    for (int t = 0; t < a.ntaps; t++) {
        const int tx = a.tap_x[t], ty = a.tap_y[t];
        const float dx = (float) tx - fx, dy = (float) ty - fy;
        const float d = sqrtf(__fmaf_rn(dx, dx, dy * dy));
        if ((a.tap_flags[t] & 1) && !(d < a.radius))
            continue;
        const float w = plb_lut1d(a.lut, a.lut_entries, d * 1.0f / a.radius);
        wsum += w;
        const float4 c = plb_load_rgba16hf(a.src, plb_clampi(jx + tx, 0, wm1), plb_clampi(jy + ty, 0, hm1));
        c0 = __fmaf_rn(w, c.x, c0);
        c1 = __fmaf_rn(w, c.y, c1);
        c2 = __fmaf_rn(w, c.z, c2);
        c3 = __fmaf_rn(w, c.w, c3);
    }
    const float norm = a.scale / wsum;
    out[0] = (a.comp_mask & 1) ? norm * c0 : 0.0f;
    out[1] = (a.comp_mask & 2) ? norm * c1 : 0.0f;
    out[2] = (a.comp_mask & 4) ? norm * c2 : 0.0f;
    out[3] = (a.comp_mask & 8) ? norm * c3 : 1.0f;
}

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
        const float w = plb_lut1d(a.lut, a.lut_entries, d * 1.0f / a.radius); // sampling.c:529, a division
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
template <bool HALF4>
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
        if (HALF4)
            polar_pixel_half4(a, gx, gy, prim);
        else
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
                    const float w = host_lut1d(p->host_lut, p->lut_entries, d * 1.0f / p->radius);
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
            // centred phases for every component count, left / top-left sited phases for chroma
            // pairs; anything else (-1) takes the generic kernel below
            const int r = (xm | ym) ? plb_up2_launch_b(gpu, &args, nc, xm, ym)
                        : nc <= 2 ? plb_up2_launch_a(gpu, &args, nc, xm, ym)
                                  : plb_up2_launch_c(gpu, &args, nc, xm, ym);
            if (r >= 0) {
                g->launches++;
                return r == 1;
            }
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
    if (a.src.kind == PLB_FK_HALF && a.src.comps == 4 && !a.src.ptr2 && a.src.addr == 0 && !(a.antiring > 0.0f))
        k_polar_generic<true><<<grid, block, 0, STREAM(gpu)>>>(a);
    else
        k_polar_generic<false><<<grid, block, 0, STREAM(gpu)>>>(a);
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
