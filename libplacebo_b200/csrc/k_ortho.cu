// pl_shader_sample_ortho2 (src/shaders/sampling.c:950-1104): one 1-D pass of a separable
// scaler. Weights come from the 256-row LUT (row = phase, bilinear in phase — the reference
// samples an rgba32f texture LINEARly, sh_lut lut.c:285-295), taps from the source texture
// through linear fetches; radius == radius_zero filters use the linear-sampling trick
// (sampling.c:919-936, 1062-1088: two taps per fetch). Same per-pixel tail as every sampler.
//
// Tap fetches: the coordinate of a tap is a texel centre along the filtered axis by construction;
// when the other axis maps 1:1 onto the whole texture (every pass the renderer's separated
// scaling issues, renderer.c:745-775) it is a texel centre there too, and a LINEAR fetch at a
// texel centre returns the texel (texture units quantise the weights to 1/256). Such passes
// (`aligned`) fetch the nearest texel; with the linear-sampling trick (fractional position along
// the filtered axis) they interpolate along that axis only. The oracle follows the same rule
// (orc_sample_ortho).
//
// CTA = 32x8 threads over a 32x32 output tile (4 rows per thread); the N <= 2*ceil(R) taps of a
// warp walk one row segment (horizontal pass) or N rows (vertical pass), so every fetch of a
// warp is a contiguous, L1-resident segment. Bound: HBM for small N (lanczos N = 6: 6 fetches
// + ~20 FMA per output), i.e. the source is read once from DRAM and N times from L1.
#include "pixel_core.cuh"

struct ortho_args {
    plb_tex_desc src;
    plb_coord cc;
    const float *lut;
    int lut_entries, row_size, row_stride;
    int use_linear;
    int aligned;     // the unfiltered axis maps 1:1 onto the whole texture (texel centres)
    float antiring, ratio, scale;
    uint8_t comp_mask;
    int pass;
    plb_tail_args tail;
};

PLB_DEV float plb_fractf(float x) { return x - floorf(x); }

PLB_DEV void ortho_weights(const ortho_args &a, int quad, float fcoord, float ws[4])
{
    const float fpos = plb_clamp(fcoord, 0.0f, 1.0f) * (float) (a.lut_entries - 1);
    const float fbase = floorf(fpos), fceil = ceilf(fpos);
    const float fc = fpos - fbase;
    const float4 r0 = __ldg((const float4 *) (a.lut + (size_t) fbase * a.row_stride + 4 * quad));
    const float4 r1 = __ldg((const float4 *) (a.lut + (size_t) fceil * a.row_stride + 4 * quad));
    ws[0] = plb_mix(r0.x, r1.x, fc);
    ws[1] = plb_mix(r0.y, r1.y, fc);
    ws[2] = plb_mix(r0.z, r1.z, fc);
    ws[3] = plb_mix(r0.w, r1.w, fc);
}

// Tap fetch. FETCH 0: bilinear (general position); 1: nearest texel (aligned pass, see the
// header); 2: interpolation along the filtered axis only (aligned pass with the linear-sampling
// trick). HALF4: rgba16hf source (the image FBO), fetched without the run-time format dispatch —
// the same texels and the same arithmetic as the general functions.
template <int FETCH, bool HALF4>
PLB_DEV float4 ortho_fetch(const ortho_args &a, float tu, float tv)
{
    if (!HALF4) {
        return FETCH == 0 ? plb_tex_linear(a.src, tu, tv)
             : FETCH == 2 ? plb_tex_linear_axis(a.src, tu, tv, a.pass == 1 ? 0 : 1)
             : plb_tex_nearest(a.src, tu, tv);
    }
    const int w = a.src.w, h = a.src.h, m = a.src.addr;
    if (FETCH == 1) {
        const int x = (int) floorf(tu * (float) w), y = (int) floorf(tv * (float) h);
        return plb_load_rgba16hf(a.src, plb_wrap_index(x, w, m), plb_wrap_index(y, h, m));
    }
    if (FETCH == 2) {
        float4 p, q;
        float f;
        if (a.pass == 1) {
            const float tx = tu * (float) w - 0.5f, bx = floorf(tx);
            const int y = plb_wrap_index((int) floorf(tv * (float) h), h, m);
            f = tx - bx;
            p = plb_load_rgba16hf(a.src, plb_wrap_index((int) bx, w, m), y);
            q = plb_load_rgba16hf(a.src, plb_wrap_index((int) bx + 1, w, m), y);
        } else {
            const float ty = tv * (float) h - 0.5f, by = floorf(ty);
            const int x = plb_wrap_index((int) floorf(tu * (float) w), w, m);
            f = ty - by;
            p = plb_load_rgba16hf(a.src, x, plb_wrap_index((int) by, h, m));
            q = plb_load_rgba16hf(a.src, x, plb_wrap_index((int) by + 1, h, m));
        }
        return make_float4(plb_mix(p.x, q.x, f), plb_mix(p.y, q.y, f), plb_mix(p.z, q.z, f), plb_mix(p.w, q.w, f));
    }
    const float tx = tu * (float) w - 0.5f, ty = tv * (float) h - 0.5f;
    const float bx = floorf(tx), by = floorf(ty);
    const float fx = tx - bx, fy = ty - by;
    const int x0 = plb_wrap_index((int) bx, w, m), x1 = plb_wrap_index((int) bx + 1, w, m);
    const int y0 = plb_wrap_index((int) by, h, m), y1 = plb_wrap_index((int) by + 1, h, m);
    const float4 p = plb_load_rgba16hf(a.src, x0, y0), q = plb_load_rgba16hf(a.src, x1, y0);
    const float4 r = plb_load_rgba16hf(a.src, x0, y1), t = plb_load_rgba16hf(a.src, x1, y1);
    return make_float4(plb_mix(plb_mix(p.x, q.x, fx), plb_mix(r.x, t.x, fx), fy),
                       plb_mix(plb_mix(p.y, q.y, fx), plb_mix(r.y, t.y, fx), fy),
                       plb_mix(plb_mix(p.z, q.z, fx), plb_mix(r.z, t.z, fx), fy),
                       plb_mix(plb_mix(p.w, q.w, fx), plb_mix(r.w, t.w, fx), fy));
}

template <int FETCH, bool HALF4>
PLB_DEV void ortho_pixel(const ortho_args &a, int gx, int gy, float out[4])
{
    const int N = a.row_size;
    const float dirx = a.pass == 1 ? 1.0f : 0.0f, diry = a.pass == 0 ? 1.0f : 0.0f;
    const bool use_ar = a.antiring > 0.0f && a.ratio > 1.0f && !a.use_linear;
    float posx, posy;
    plb_coord_pos(a.cc, gx, gy, posx, posy);
    const float f2x = plb_fractf(posx * a.cc.sizex - 0.5f), f2y = plb_fractf(posy * a.cc.sizey - 0.5f);
    const float fcoord = f2x * dirx + f2y * diry;
    const float ptx = a.cc.ptx * dirx, pty = a.cc.pty * diry;
    const float off0 = (float) (N / 2 - 1);
    const float basex = posx - fcoord * ptx - ptx * off0;
    const float basey = posy - fcoord * pty - pty * off0;
    float ca[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, lo[4] = {1e9f, 1e9f, 1e9f, 1e9f};
    float ws[4] = {0, 0, 0, 0};
    // aligned pass over an rgba16hf image: the texel index along the unfiltered axis does not
    // depend on the tap (pt = 0 there, so tu resp. tv is the base coordinate itself)
    const bool fast = HALF4 && FETCH != 0;
    const int sw = a.src.w, sh = a.src.h, mode = a.src.addr;
    int fixed = 0;
    if (fast)
        fixed = a.pass == 0 ? plb_wrap_index((int) floorf(basex * (float) sw), sw, mode)
                            : plb_wrap_index((int) floorf(basey * (float) sh), sh, mode);
    // taps in groups of four (one LUT quad): static indices into ws[] keep it in registers
    for (int n0 = 0; n0 < N; n0 += 4) {
        ortho_weights(a, n0 / 4, fcoord, ws);
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + j;
            if (n >= N || (a.use_linear && (j & 1)))
                continue;
            float off = (float) n;
            if (a.use_linear)
                off += ws[(j + 1) & 3];
            const float tu = basex + ptx * off, tv = basey + pty * off;
            float4 c4;
            if (fast) {
                const int lim = a.pass == 0 ? sh : sw;
                const float t = a.pass == 0 ? tv * (float) sh : tu * (float) sw;
                if (FETCH == 1) {
                    const int i = plb_wrap_index((int) floorf(t), lim, mode);
                    c4 = a.pass == 0 ? plb_load_rgba16hf(a.src, fixed, i) : plb_load_rgba16hf(a.src, i, fixed);
                } else {
                    const float tt = t - 0.5f, b = floorf(tt), f = tt - b;
                    const int i0 = plb_wrap_index((int) b, lim, mode), i1 = plb_wrap_index((int) b + 1, lim, mode);
                    const float4 p = a.pass == 0 ? plb_load_rgba16hf(a.src, fixed, i0) : plb_load_rgba16hf(a.src, i0, fixed);
                    const float4 q = a.pass == 0 ? plb_load_rgba16hf(a.src, fixed, i1) : plb_load_rgba16hf(a.src, i1, fixed);
                    c4 = make_float4(plb_mix(p.x, q.x, f), plb_mix(p.y, q.y, f), plb_mix(p.z, q.z, f), plb_mix(p.w, q.w, f));
                }
            } else {
                c4 = ortho_fetch<FETCH, HALF4>(a, tu, tv);
            }
            const float c[4] = {c4.x, c4.y, c4.z, c4.w};
            if (use_ar && (n == N / 2 - 1 || n == N / 2)) {
                #pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < a.src.comps) {
                        lo[k] = fminf(lo[k], c[k]);
                        hi[k] = fmaxf(hi[k], c[k]);
                    }
                }
            }
            const float w = ws[j];
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                if (HALF4 || k < a.src.comps)
                    ca[k] = __fmaf_rn(w, c[k], ca[k]);
            }
        }
    }
    out[0] = out[1] = out[2] = 0.0f;
    out[3] = 1.0f;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!(a.comp_mask & (1 << k)) || k >= a.src.comps)
            continue;
        float v = ca[k];
        if (use_ar)
            v = plb_mix(v, plb_clamp(v, lo[k], hi[k]), a.antiring);
        out[k] = a.scale * v;
    }
}

template <int FETCH, bool HALF4>
__global__ void __launch_bounds__(256)
k_ortho(const __grid_constant__ ortho_args a)
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
        ortho_pixel<FETCH, HALF4>(a, gx, gy, prim);
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

// The same pass for a one-component half-float plane, aligned taps, no tail (the
// contrast-recovery low-pass, renderer.c:2133-2150): the value equals ortho_pixel()'s.
PLB_DEV float ortho_h1_texel(const ortho_args &a, int x, int y)
{
    return __half2float(*(const __half *) ((const char *) a.src.ptr + (size_t) y * a.src.pitch + 2 * (size_t) x));
}

template <int PASS, bool LIN>
__global__ void __launch_bounds__(256)
k_ortho_h1(const __grid_constant__ ortho_args a)
{
    const int gx = blockIdx.x * 32 + threadIdx.x;
    if (gx >= a.tail.out_w)
        return;
    const int N = a.row_size;
    const int sw = a.src.w, sh = a.src.h, mode = a.src.addr;
    #pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const int gy = blockIdx.y * 32 + threadIdx.y + 8 * r;
        if (gy >= a.tail.out_h)
            break;
        float posx, posy;
        plb_coord_pos(a.cc, gx, gy, posx, posy);
        const float f2x = plb_fractf(posx * a.cc.sizex - 0.5f), f2y = plb_fractf(posy * a.cc.sizey - 0.5f);
        const float dirx = PASS == 1 ? 1.0f : 0.0f, diry = PASS == 0 ? 1.0f : 0.0f;
        const float fcoord = f2x * dirx + f2y * diry;
        const float ptx = a.cc.ptx * dirx, pty = a.cc.pty * diry;
        const float off0 = (float) (N / 2 - 1);
        const float basex = posx - fcoord * ptx - ptx * off0;
        const float basey = posy - fcoord * pty - pty * off0;
        // the coordinate of the unfiltered axis does not depend on the tap (pt = 0)
        const int fixed = PASS == 0 ? plb_wrap_index((int) floorf(basex * (float) sw), sw, mode)
                                    : plb_wrap_index((int) floorf(basey * (float) sh), sh, mode);
        float ca = 0.0f;
        float ws[4] = {0, 0, 0, 0};
        for (int n0 = 0; n0 < N; n0 += 4) {
            ortho_weights(a, n0 / 4, fcoord, ws);
            #pragma unroll
            for (int j = 0; j < 4; j += LIN ? 2 : 1) {
                if (n0 + j >= N)
                    break;
                float off = (float) (n0 + j);
                if (LIN)
                    off += ws[(j + 1) & 3];
                const float t = PASS == 0 ? (basey + pty * off) * (float) sh : (basex + ptx * off) * (float) sw;
                const int lim = PASS == 0 ? sh : sw;
                float c;
                if (LIN) {
                    const float tt = t - 0.5f, b = floorf(tt), f = tt - b;
                    const int i0 = plb_wrap_index((int) b, lim, mode), i1 = plb_wrap_index((int) b + 1, lim, mode);
                    const float c0 = PASS == 0 ? ortho_h1_texel(a, fixed, i0) : ortho_h1_texel(a, i0, fixed);
                    const float c1 = PASS == 0 ? ortho_h1_texel(a, fixed, i1) : ortho_h1_texel(a, i1, fixed);
                    c = plb_mix(c0, c1, f);
                } else {
                    const int i = plb_wrap_index((int) floorf(t), lim, mode);
                    c = PASS == 0 ? ortho_h1_texel(a, fixed, i) : ortho_h1_texel(a, i, fixed);
                }
                ca = __fmaf_rn(ws[j], c, ca);
            }
        }
        *(__half *) ((char *) a.tail.dst.ptr + (size_t) (a.tail.dst_y0 + gy) * a.tail.dst.pitch +
                     2 * (size_t) (a.tail.dst_x0 + gx)) = __float2half_rn(a.scale * ca);
    }
}

#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)
void plb_fill_tail(plb_tail_args *t, const struct plb_pixel_launch *p);
PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_ortho)

extern "C" bool plb_launch_ortho(pl_gpu gpu, const struct plb_ortho_launch *p)
{
    struct plb_gpu *g = plb_gpu_priv(gpu);
    const struct plb_pixel_launch *tl = &p->tail;
    if (p->row_size < 2 || p->row_size > 64 || p->row_stride % 4) {
        PLB_ERR(gpu->log, "ortho sampling: unsupported filter size %d", p->row_size);
        return false;
    }
    bool has_peak = false;
    for (int i = 0; i < tl->prog.n; i++)
        has_peak |= tl->prog.ops[i].kind == PLB_OP_PEAK;
    const int wg_w = PLB_DEF(tl->wg_w, 16), wg_h = PLB_DEF(tl->wg_h, 16);
    if (has_peak && (32 % wg_w || 32 % wg_h)) {
        PLB_ERR(gpu->log, "peak detection: unsupported work group size %dx%d", wg_w, wg_h);
        return false;
    }
    static ortho_args a;
    memset(&a, 0, sizeof(a));
    a.src = p->src;
    a.cc = plb_coord_init(tl->coord_mode, p->src, p->rect, tl->out_w, tl->out_h);
    a.lut = p->lut;
    a.lut_entries = p->lut_entries;
    a.row_size = p->row_size;
    a.row_stride = p->row_stride;
    a.use_linear = p->use_linear;
    a.antiring = p->antiring;
    a.ratio = p->ratio;
    a.scale = p->scale;
    a.comp_mask = p->comp_mask;
    a.pass = p->pass;
    plb_fill_tail(&a.tail, tl);
    a.tail.wg_w = wg_w;
    a.tail.wg_h = wg_h;
    memcpy(a.tail.src_map, p->src_map, sizeof(p->src_map));
    // taps on texel centres of both axes (see the header): nearest fetches
    const bool aligned = tl->coord_mode != PLB_COORD_LITERAL &&
        (p->pass == 0 ? (p->rect.x0 == 0.0f && p->rect.x1 == (float) p->src.w && tl->out_w == p->src.w)
                      : (p->rect.y0 == 0.0f && p->rect.y1 == (float) p->src.h && tl->out_h == p->src.h));
    a.aligned = aligned;
    dim3 block(32, 8);
    dim3 grid((tl->out_w + 31) / 32, (tl->out_h + 31) / 32);
    const bool use_ar = p->antiring > 0.0f && p->ratio > 1.0f && !p->use_linear;
    if (aligned && !use_ar && !has_peak && tl->prog.n == 0 && tl->nplanes == 0 && tl->dst.ptr &&
        p->src.kind == PLB_FK_HALF && p->src.comps == 1 && !p->src.ptr2 &&
        tl->dst.kind == PLB_FK_HALF && tl->dst.comps == 1 && !tl->dst.ptr2 &&
        (p->comp_mask & 1) && p->src_map[0] == 0) {
        if (p->pass == 0 && p->use_linear)
            k_ortho_h1<0, true><<<grid, block, 0, STREAM(gpu)>>>(a);
        else if (p->pass == 0)
            k_ortho_h1<0, false><<<grid, block, 0, STREAM(gpu)>>>(a);
        else if (p->use_linear)
            k_ortho_h1<1, true><<<grid, block, 0, STREAM(gpu)>>>(a);
        else
            k_ortho_h1<1, false><<<grid, block, 0, STREAM(gpu)>>>(a);
        g->launches++;
        return plb_cuda_check(gpu, cudaGetLastError(), "k_ortho_h1");
    }
    if (has_peak) {
        grid.x = ((tl->out_w + wg_w - 1) / wg_w * wg_w + 31) / 32;
        grid.y = ((tl->out_h + wg_h - 1) / wg_h * wg_h + 31) / 32;
    }
    const int fetch = !a.aligned ? 0 : a.use_linear ? 2 : 1;
    const bool half4 = p->src.kind == PLB_FK_HALF && p->src.comps == 4 && !p->src.ptr2;
#define ORTHO(F, H) k_ortho<F, H><<<grid, block, 0, STREAM(gpu)>>>(a)
    if (half4) {
        if (fetch == 0) ORTHO(0, true); else if (fetch == 1) ORTHO(1, true); else ORTHO(2, true);
    } else {
        if (fetch == 0) ORTHO(0, false); else if (fetch == 1) ORTHO(1, false); else ORTHO(2, false);
    }
#undef ORTHO
    g->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_ortho");
}
