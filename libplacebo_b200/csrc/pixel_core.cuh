// The per-pixel tail shared by every sampling kernel (polar, ortho, plain pixel pass):
//   1. merge extra planes into the colour (renderer.c:1788-1897: "color[map] = tmp[k]"),
//      each plane being sampled directly, debanded (sampling.c:183-275) or grain-synthesised
//   2. run the recorded colour ops (color_ops.cuh)
//   3. fold the pixel into the peak-detection statistics of its work group
//      (colorspace.c:1290-1348), using warp match/reduce + shared-memory atomics
//   4. convert and store to the target texture
#pragma once

#include <cstddef>

#include "color_ops.cuh"

struct plb_tail_args {
    int nplanes;
    plb_plane_src planes[4];
    plb_coord plane_cc[4];
    float init[4];
    int8_t src_map[4];     // colour index for component k of the primary sampler (-1 none)
    plb_tex_desc dst;      // ptr == NULL: statistics only
    int dst_x0, dst_y0;
    int out_w, out_h;
    uint32_t frame_index;
    int wg_w, wg_h;        // work-group geometry for PLB_OP_PEAK
    plb_color_prog prog;
    // after .prog: staged to shared memory only when a reshape op is present
    plb_dovi_tab dovi;
};

// Per-CTA shared-memory copy of the launch description (everything but .prog, which lives in
// plb_s_prog). Every kernel calls plb_tail_stage() with its parameter first, then a barrier.
static __shared__ plb_tail_args plb_s_tail;

PLB_DEV void plb_tail_stage(const plb_tail_args &a)
{
    const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
    const int nt = blockDim.x * blockDim.y * blockDim.z;
    const uint32_t *src = (const uint32_t *) &a;
    uint32_t *dst = (uint32_t *) &plb_s_tail;
    constexpr int head = (int) (offsetof(plb_tail_args, prog) / 4);
    for (int i = tid; i < head; i += nt)
        dst[i] = src[i];
    // only the used part of the program
    const uint32_t *psrc = (const uint32_t *) &a.prog;
    uint32_t *pdst = (uint32_t *) &plb_s_prog;
    const int pwords = (int) ((offsetof(plb_color_prog, ops) + (size_t) a.prog.n * sizeof(plb_color_op)) / 4);
    for (int i = tid; i < pwords; i += nt)
        pdst[i] = psrc[i];
    if (tid == 0)
        plb_s_tail.prog.n = a.prog.n;
    bool dovi = false;
    for (int i = 0; i < a.prog.n; i++)
        dovi |= a.prog.ops[i].kind == PLB_OP_DOVI_RESHAPE;
    if (dovi) {
        const uint32_t *dsrc = (const uint32_t *) &a.dovi;
        uint32_t *ddst = (uint32_t *) &plb_s_dovi;
        for (int i = tid; i < (int) (sizeof(plb_dovi_tab) / 4); i += nt)
            ddst[i] = dsrc[i];
    }
}

#define PLB_MAX_WG_PER_CTA 16
#define PLB_HIST_BINS 64
#define PLB_PEAK_SLICES 12

struct plb_peak_smem {
    uint32_t sum[PLB_MAX_WG_PER_CTA], mx[PLB_MAX_WG_PER_CTA], black[PLB_MAX_WG_PER_CTA];
    uint32_t hist[PLB_MAX_WG_PER_CTA][PLB_HIST_BINS];
};

// pl_shader_deband for one plane at output pixel (gx, gy); c = texel at pos (already loaded)
// FAST: layout class of pl.tex (plb_tex_fast_kind), 0 = any
template <int FAST = 0>
PLB_DEV void plb_deband(const plb_plane_src &pl, const plb_coord &cc, float posx, float posy,
                        int gx, int gy, uint32_t frame_index, float color[4])
{
    const uint8_t mask = pl.comp_mask & ~0x8u;
    if (!mask)
        return;
    uint3 state = make_uint3((uint32_t) gx, (uint32_t) gy, frame_index);
    float res[3] = { color[0], color[1], color[2] };
    for (int i = 1; i <= pl.iterations; i++) {
        const float3 rnd = plb_prng(state);
        const float dist = rnd.x * ((float) i * pl.radius);
        const float dir = rnd.y * 6.283185f;
        float sn, cs;
        plm_sincos_fast(dir, &sn, &cs);
        const float dx = dist * cs, dy = dist * sn;
        float avg[3] = {0.0f, 0.0f, 0.0f};
        #pragma unroll
        for (int t = 0; t < 4; t++) {
            const float sx = (t == 0 || t == 3) ? 1.0f : -1.0f, sy = t < 2 ? 1.0f : -1.0f;
            const float4 c = plb_tex_nearest_t<FAST>(pl.tex, posx + cc.ptx * (sx * dx),
                                                     posy + cc.pty * (sy * dy));
            avg[0] += c.x; avg[1] += c.y; avg[2] += c.z;
        }
        const float bound = pl.threshold / (float) i;
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            if (!(mask & (1 << k)))
                continue;
            const float a = avg[k] * 0.25f;
            const float diff = fabsf(res[k] - a);
            res[k] = diff > bound ? res[k] : a;
        }
    }
    if (pl.grain_str > 0.0f) {
        const float3 rnd3 = plb_prng(state);
        const float rnd[3] = { rnd3.x, rnd3.y, rnd3.z };
        int idx = 0;
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            if (!(mask & (1 << k)))
                continue;
            const float bound = pl.neutral[idx] / pl.scale;
            const float strength = fminf(fabsf(res[k] - bound), pl.grain_str);
            res[k] = __fmaf_rn(strength, rnd[idx] - 0.5f, res[k]);
            idx++;
        }
    }
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        if (mask & (1 << k))
            color[k] = res[k];
    }
}

// pl_shader_fg_av1 (film_grain_av1.c:827-1013): grain = LUT[block offset + local id], blended
// over the 2-pixel block overlap, scaled by a luma-dependent LUT, added and clamped.
PLB_DEV float plb_av1_fetch(const plb_av1_plane &g, int i, uint32_t data, int off, int lx, int ly,
                            int bw, int bh)
{
    const int dx = (off & 8) ? 1 : 0, dy = (off & 16) ? 1 : 0; // OFFSET_L = 8, OFFSET_T = 16
    const int ox = (g.sub_x ? 1 : 2) * (int) ((data >> (off + 4)) & 0xFu);
    const int oy = (g.sub_y ? 1 : 2) * (int) ((data >> off) & 0xFu);
    const int px = ox + lx + bw * dx, py = oy + ly + bh * dy;
    return __ldg(g.grain[i] + py * g.grain_w[i] + px);
}

PLB_DEV void plb_grain_av1(const plb_plane_src &pl, int gx, int gy, float color[4])
{
    const plb_av1_plane &g = pl.av1;
    const int bw = 32 >> g.sub_x, bh = 32 >> g.sub_y;
    const int bx = gx / bw, by = gy / bh;
    const int lx = gx - bw * bx, ly = gy - bh * by;
    const uint32_t data = __ldg(g.offsets + by * g.off_w + bx);
    const float4 t = plb_load_clamped(pl.tex, gx, gy);
    color[0] = g.tex_scale * t.x;
    color[1] = g.tex_scale * t.y;
    color[2] = g.tex_scale * t.z;
    color[3] = g.tex_scale * t.w;
    float avg_luma = 0.0f;
    if (g.luma_from_comp >= 0) {
        avg_luma = color[g.luma_from_comp];
    } else if (g.luma_from_comp == -1) {
        const float4 l = plb_load_clamped(pl.luma, gx << g.sub_x, gy << g.sub_y);
        const float lv[4] = { l.x, l.y, l.z, l.w };
        avg_luma = lv[g.luma_comp] * g.tex_scale;
    }
    const float wx0 = g.sub_x ? 23.0f / 32.0f : 27.0f / 32.0f, wx1 = g.sub_x ? 22.0f / 32.0f : 17.0f / 32.0f;
    const float wy0 = g.sub_y ? 23.0f / 32.0f : 27.0f / 32.0f, wy1 = g.sub_y ? 22.0f / 32.0f : 17.0f / 32.0f;
    for (int i = 0; i < g.ncomp; i++) {
        if (!g.scaling[i])
            continue;
        float grain = plb_av1_fetch(g, i, data, 0, lx, ly, bw, bh);
        if (g.overlap) {
            const bool ovx = bx > 0 && lx < (2 >> g.sub_x);
            if (ovx) {
                const float a = lx == 1 ? wx1 : wx0, b = lx == 1 ? wx0 : wx1;
                const float val = plb_av1_fetch(g, i, data, 8, lx, ly, bw, bh);
                grain = __fmaf_rn(grain, b, val * a); // dot(vec2(val, grain), w)
            }
            if (by > 0 && ly < (2 >> g.sub_y)) {
                const float a = ly == 1 ? wy1 : wy0, b = ly == 1 ? wy0 : wy1;
                float val;
                if (ovx) {
                    const float a2 = lx == 1 ? wx1 : wx0, b2 = lx == 1 ? wx0 : wx1;
                    const float tmp = plb_av1_fetch(g, i, data, 24, lx, ly, bw, bh);
                    val = plb_av1_fetch(g, i, data, 16, lx, ly, bw, bh);
                    val = __fmaf_rn(val, b2, tmp * a2);
                } else {
                    val = plb_av1_fetch(g, i, data, 16, lx, ly, bw, bh);
                }
                grain = __fmaf_rn(grain, b, val * a);
            }
            grain = plb_clamp(grain, g.grain_min, g.grain_max);
        }
        if (g.channel[i] == 0) { // PL_CHANNEL_Y
            const float s = plb_lut1d(g.scaling[i], 256, color[i]);
            color[i] = __fmaf_rn(s, grain, color[i]);
            color[i] = plb_clamp(color[i], g.min_value, g.max_luma);
        } else {
            float val = avg_luma;
            if (!g.chroma_from_luma) {
                val = __fmaf_rn(color[i], g.mult[i][1], val * g.mult[i][0]);
                val += g.offset[i];
            }
            const float s = plb_lut1d(g.scaling[i], 256, val);
            color[i] = __fmaf_rn(s, grain, color[i]);
            color[i] = plb_clamp(color[i], g.min_value, g.max_chroma);
        }
    }
}

// pl_shader_fg_h274 (film_grain_h274.c:145-315). The reference runs 8x8 work groups: the block
// mean of each component is a subgroupAdd over the (two) 32-wide subgroups of the group, merged
// through an integer atomicAdd of int(sum * 1e7). GLSL leaves the order of subgroupAdd open; here
// (and in the oracle) it is DEFINED as ascending invocation index. Every thread evaluates the
// mean of its own block from the (L1-resident) texels, which needs no cross-thread exchange.
// A real function on purpose: the grain pass is rare, and inlined into the common pixel tail
// it cost the HDR colour pass 14 % (instruction-cache footprint, register allocation).
static __device__ __noinline__ void plb_grain_h274(const plb_plane_src &pl, int gx, int gy, float color[4])
{
    const plb_h274_plane &g = *pl.h274;
    const float4 t = plb_load_clamped(pl.tex, gx, gy);
    color[0] = g.tex_scale * t.x;
    color[1] = g.tex_scale * t.y;
    color[2] = g.tex_scale * t.z;
    color[3] = g.tex_scale * t.w;
    const int bx = gx >> 3, by = gy >> 3;
    // pcg3d of (seed, block pair)
    uint32_t px = g.seed, py = (uint32_t) (bx >> 1), pz = (uint32_t) (by >> 1);
    px = px * 1664525u + 1013904223u; py = py * 1664525u + 1013904223u; pz = pz * 1664525u + 1013904223u;
    px += py * pz; py += pz * px; pz += px * py;
    px ^= px >> 16; py ^= py >> 16; pz ^= pz >> 16;
    px += py * pz; py += pz * px; pz += px * py;
    const uint32_t pcg[3] = { px, py, pz };
    for (int i = 0; i < g.nent; i++) {
        const int c = g.ent[i].channel;
        // block mean: two subgroups of 32 invocations (4 rows each), ascending order
        int isum = 0;
        #pragma unroll 1
        for (int half = 0; half < 2; half++) {
            float s = 0.0f;
            for (int r = 0; r < 4; r++) {
                for (int col = 0; col < 8; col++) {
                    const int x = 8 * bx + col, y = 8 * by + 4 * half + r;
                    float v = 0.0f; // texelFetch outside the image: robust-access zero
                    if (x < pl.tex.w && y < pl.tex.h) {
                        const float4 q = plb_load(pl.tex, x, y);
                        v = g.tex_scale * (c == 0 ? q.x : c == 1 ? q.y : c == 2 ? q.z : q.w);
                    }
                    s += v / 64.0f;
                }
            }
            isum += (int) (s * 10000000.0f);
        }
        const float avg = (float) isum / 10000000.0f;
        uint32_t val = 0u;
        const plb_h274_interval *iv = g.iv + g.ent[i].first;
        for (int k = g.ent[i].count - 1; k >= 0; k--) {
            // the GLSL is an if / else-if chain in ascending k: the first match wins
            if (avg >= iv[k].lo && avg <= iv[k].hi)
                val = iv[k].value;
        }
        uint32_t ox = (val & 0xFF00u) >> 2, oy = (val & 0xFFu) << 6;
        float scale = g.scale_factor * (float) (int) (val >> 16);
        const uint32_t rnd = pcg[c];
        ox += (rnd >> 16) % 52u;
        oy += (rnd & 0xFFFFu) % 56u;
        ox &= 0xFFFCu;
        oy &= 0xFFF8u;
        if ((rnd & 1u) == 1u)
            scale = -scale;
        ox += 8u * (uint32_t) (bx & 1) + (uint32_t) (gx & 7);
        oy += 8u * (uint32_t) (by & 1) + (uint32_t) (gy & 7);
        const float grain = __ldg(g.db + (size_t) oy * (13 * 64) + ox);
        color[c] = __fmaf_rn(scale, grain, color[c]);
    }
}

// Built-in fast scalers, restated operation by operation from the GLSL (sampling.c:319-478);
// `u, v` = the interpolated normalised position `pos`, pt = 1 / texture size. Real functions:
// they are rare and must not grow the common pixel tail.
PLB_DEV float plb_fract(float x) { return x - floorf(x); }

static __device__ __noinline__ float4 plb_sample_fast(const plb_plane_src &pl, const plb_coord &cc,
                                                      float u, float v)
{
    const float sx = (float) pl.tex.w, sy = (float) pl.tex.h;
    const float ptx = cc.ptx, pty = cc.pty; // (float) (1.0 / size), as the host passes `pt`
    if (pl.kind == PLB_PLANE_HERMITE) {
        const float fx = plb_fract(u * sx + 0.5f), fy = plb_fract(v * sy + 0.5f);
        // smoothstep(0, 1, t) = t * t * (3 - 2 t) on the clamped argument
        const float tx = plb_clamp(fx, 0.0f, 1.0f), ty = plb_clamp(fy, 0.0f, 1.0f);
        const float smx = tx * tx * (3.0f - 2.0f * tx), smy = ty * ty * (3.0f - 2.0f * ty);
        return plb_tex_linear(pl.tex, u + ptx * (smx - fx), v + pty * (smy - fy));
    }
    if (pl.kind == PLB_PLANE_OVERSAMPLE) {
        const float fx = plb_fract(u * sx - 0.5f), fy = plb_fract(v * sy - 0.5f);
        float cx = plb_clamp((fx - 0.5f) * pl.over_rx + 0.5f, 0.0f, 1.0f);
        float cy = plb_clamp((fy - 0.5f) * pl.over_ry + 0.5f, 0.0f, 1.0f);
        if (pl.over_threshold > 0.0f) {
            const float th = pl.over_threshold;
            // mix(coeff, 0, lessThan(coeff, th)), mix(coeff, 1, greaterThan(coeff, 1 - th))
            cx = cx < th ? 0.0f : cx;
            cy = cy < th ? 0.0f : cy;
            cx = cx > 1.0f - th ? 1.0f : cx;
            cy = cy > 1.0f - th ? 1.0f : cy;
        }
        return plb_tex_linear(pl.tex, u + (cx - fx) * ptx, v + (cy - fy) * pty);
    }
    // bicubic / gaussian: weights of the four taps per axis, folded into two bilinear fetches
    float g[4], h[4]; // g = (x01, y01, x23, y23) sums, h = offsets in texels
    float offx = 0.0f, offy = 0.0f;
    if (pl.kind == PLB_PLANE_BICUBIC) {
        const float f[2] = { plb_fract(u * sx + 0.5f), plb_fract(v * sy + 0.5f) };
        #pragma unroll
        for (int a = 0; a < 2; a++) {
            const float fr = f[a], fr2 = fr * fr, inv = 1.0f - fr, inv2 = inv * inv;
            const float w0 = (1.0f / 6.0f) * inv2 * inv;
            const float w1 = (2.0f / 3.0f) - 0.5f * fr2 * (2.0f - fr);
            const float w2 = (2.0f / 3.0f) - 0.5f * inv2 * (2.0f - inv);
            const float w3 = (1.0f / 6.0f) * fr2 * fr;
            g[a] = w0 + w1;
            g[2 + a] = w2 + w3;
            h[a] = (w1 / g[a] + inv) - 2.0f;
            h[2 + a] = w3 / g[2 + a] + inv;
        }
    } else { // gaussian
        const float o[2] = { -plb_fract(u * sx + 0.5f), -plb_fract(v * sy + 0.5f) };
        offx = o[0];
        offy = o[1];
        #pragma unroll
        for (int a = 0; a < 2; a++) {
            const float off = o[a], off2 = -2.0f * off * off;
            const float w0 = plm_expf((off2 + 4.0f * off) - 2.0f);
            const float w1 = plm_expf(off2);
            const float w2 = plm_expf((off2 - 4.0f * off) - 2.0f);
            const float w3 = plm_expf((off2 - 8.0f * off) - 8.0f);
            g[a] = w0 + w1;
            g[2 + a] = w2 + w3;
            h[a] = w1 / g[a] - 1.0f;
            h[2 + a] = w3 / g[2 + a] + 1.0f;
        }
        g[0] = g[0] / (g[0] + g[2]); // explicit normalisation of the mixing weights
        g[1] = g[1] / (g[1] + g[3]);
        h[0] += offx; h[1] += offy; h[2] += offx; h[3] += offy;
    }
    const float px = u + ptx * h[0], py = v + pty * h[1], pz = u + ptx * h[2], pw = v + pty * h[3];
    const float4 c00 = plb_tex_linear(pl.tex, px, py), c01 = plb_tex_linear(pl.tex, px, pw);
    const float4 c10 = plb_tex_linear(pl.tex, pz, py), c11 = plb_tex_linear(pl.tex, pz, pw);
    float4 r;
#define MIX3(F) r.F = plb_mix(plb_mix(c11.F, c10.F, g[1]), plb_mix(c01.F, c00.F, g[1]), g[0]);
    MIX3(x) MIX3(y) MIX3(z) MIX3(w)
#undef MIX3
    return r;
}

// Directly sampled, debanded or AV1-grained plane at output pixel (gx, gy), BEFORE the plane's
// scale factor (the specialised kernels' inline form of plb_sample_plane for these kinds; r16
// planes take the compile-time-layout fetch).
template <int FAST>
PLB_DEV void plb_sample_plane_simple_t(const plb_plane_src &pl, const plb_coord &cc, int gx, int gy,
                                       uint32_t frame_index, float out[4])
{
    float posx, posy;
    plb_coord_pos(cc, gx, gy, posx, posy);
    const float4 c4 = (pl.kind == PLB_PLANE_DIRECT && pl.linear && !cc.ident) ? plb_tex_linear(pl.tex, posx, posy)
                                                                 : plb_tex_nearest_t<FAST>(pl.tex, posx, posy);
    out[0] = c4.x; out[1] = c4.y; out[2] = c4.z; out[3] = c4.w;
    if (pl.kind == PLB_PLANE_DEBAND)
        plb_deband<FAST>(pl, cc, posx, posy, gx, gy, frame_index, out);
    else if (pl.kind == PLB_PLANE_GRAIN_AV1) // overhanging invocations: the edge pixel's value
        plb_grain_av1(pl, min(gx, pl.tex.w - 1), min(gy, pl.tex.h - 1), out);
}

PLB_DEV void plb_sample_plane_simple(const plb_plane_src &pl, const plb_coord &cc, int gx, int gy,
                                     uint32_t frame_index, float out[4])
{
    const int fast = plb_tex_fast_kind(pl.tex); // launch-uniform
    if (fast == 1)
        plb_sample_plane_simple_t<1>(pl, cc, gx, gy, frame_index, out);
    else if (fast == 2)
        plb_sample_plane_simple_t<2>(pl, cc, gx, gy, frame_index, out);
    else
        plb_sample_plane_simple_t<0>(pl, cc, gx, gy, frame_index, out);
}

// Sample plane `pl` at output pixel (gx, gy) -> vec4 (already multiplied by its scale)
PLB_DEV void plb_sample_plane(const plb_plane_src &pl, const plb_coord &cc, int gx, int gy,
                              uint32_t frame_index, float out[4])
{
    float posx, posy;
    plb_coord_pos(cc, gx, gy, posx, posy);
    float4 c4;
    if (pl.kind >= PLB_PLANE_BICUBIC)
        c4 = plb_sample_fast(pl, cc, posx, posy);
    else if (pl.kind == PLB_PLANE_DIRECT && pl.linear && !cc.ident)
        c4 = plb_tex_linear(pl.tex, posx, posy);
    else
        c4 = plb_tex_nearest(pl.tex, posx, posy);
    out[0] = c4.x; out[1] = c4.y; out[2] = c4.z; out[3] = c4.w;
    if (pl.kind == PLB_PLANE_DEBAND)
        plb_deband(pl, cc, posx, posy, gx, gy, frame_index, out);
    if (pl.kind == PLB_PLANE_GRAIN_H274) {
        plb_grain_h274(pl, gx, gy, out);
        #pragma unroll
        for (int k = 0; k < 4; k++)
            out[k] *= pl.scale;
        return;
    }
    if (pl.kind == PLB_PLANE_GRAIN_AV1) {
        // invocations beyond the plane (work-group overhang, counted by peak detection) take the
        // edge pixel's value, as if the grain pass had gone through an FBO sampled with clamping
        plb_grain_av1(pl, min(gx, pl.tex.w - 1), min(gy, pl.tex.h - 1), out);
        #pragma unroll
        for (int k = 0; k < 4; k++)
            out[k] *= pl.scale; // the renderer's "tmp = vec4(scale) * sub()"
        return;
    }
    #pragma unroll
    for (int k = 0; k < 4; k++)
        out[k] *= pl.scale;
}

// Steps 1-2 + the pixel's PQ luma for step 3. `prim` = output of the primary sampler (or NULL)
// Not inlined on purpose (one copy per translation unit; see plb_run_color_ops).
static __device__ __noinline__ bool plb_tail_color_slow(int gx, int gy,
                                                 const float *prim, float color[4],
                                                 uint32_t *ypq, const plb_op_peak **pk);

PLB_DEV bool plb_tail_color(int gx, int gy, const float *prim,
                            float color[4], uint32_t *ypq, const plb_op_peak **pk)
{
    const plb_tail_args &a = plb_s_tail;
    if (a.nplanes == 0 && a.prog.n == 0 && prim) {
        // pure resampling pass: nothing to merge, nothing to run
        #pragma unroll
        for (int k = 0; k < 4; k++)
            color[k] = a.init[k];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (a.src_map[k] >= 0)
                color[a.src_map[k]] = prim[k];
        }
        *pk = nullptr;
        return false;
    }
    return plb_tail_color_slow(gx, gy, prim, color, ypq, pk);
}

static __device__ __noinline__ bool plb_tail_color_slow(int gx, int gy, const float *prim,
                            float color[4], uint32_t *ypq, const plb_op_peak **pk)
{
    const plb_tail_args &a = plb_s_tail;
    #pragma unroll
    for (int k = 0; k < 4; k++)
        color[k] = a.init[k];
    if (prim) {
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (a.src_map[k] >= 0)
                color[a.src_map[k]] = prim[k];
        }
    }
    for (int i = 0; i < a.nplanes; i++) {
        float tmp[4];
        plb_sample_plane(a.planes[i], a.plane_cc[i], gx, gy, a.frame_index, tmp);
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (a.planes[i].map[k] >= 0)
                color[a.planes[i].map[k]] = tmp[k];
        }
    }
    *pk = nullptr;
    if (a.prog.n)
        plb_run_color_ops(gx, gy, color, pk, ypq);
    return *pk != nullptr;
}

// Step 3, block-cooperative. All threads of the CTA must call plb_peak_begin/finish; pixels
// are added with plb_peak_add (by any subset of lanes, so it uses a full-warp vote first).
PLB_DEV void plb_peak_begin(plb_peak_smem &s)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int i = tid; i < PLB_MAX_WG_PER_CTA; i += nt)
        s.sum[i] = s.mx[i] = s.black[i] = 0;
    for (int i = tid; i < PLB_MAX_WG_PER_CTA * PLB_HIST_BINS; i += nt)
        (&s.hist[0][0])[i] = 0;
    __syncthreads();
}

// `valid`: this lane contributes a pixel (its work group exists in the reference's grid).
// Must be called convergently by all 32 lanes of the warp.
PLB_DEV void plb_peak_add(plb_peak_smem &s, const plb_op_peak &p, bool valid, int wg_local,
                          uint32_t ypq)
{
    const unsigned full = 0xFFFFFFFFu;
    const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
    const int key = valid ? wg_local : -1;
    // usual case: the warp's 32 pixels lie in one work group (row segments of tiles whose work
    // groups are >= 32 wide): one vote instead of a match
    const int key0 = __shfl_sync(full, key, 0);
    const unsigned m = __all_sync(full, key == key0) ? full : __match_any_sync(full, key);
    if (valid) {
        const uint32_t gsum = __reduce_add_sync(m, ypq);
        const uint32_t gmax = __reduce_max_sync(m, ypq);
        const unsigned blackm = __ballot_sync(m, ypq == 0u) & m;
        if (lane == __ffs(m) - 1) {
            atomicAdd(&s.sum[wg_local], gsum);
            atomicMax(&s.mx[wg_local], gmax);
            if (p.cutoff != 0.0f)
                atomicAdd(&s.black[wg_local], (uint32_t) __popc(blackm));
        }
    }
    if (p.use_histogram) {
        int bin = (int) ypq >> (14 - 7);
        bin -= 64;
        bin = plb_clampi(bin, 0, PLB_HIST_BINS - 1);
        const int hkey = valid ? wg_local * PLB_HIST_BINS + bin : -1;
        const unsigned hm = __match_any_sync(full, hkey);
        if (valid && lane == __ffs(hm) - 1)
            atomicAdd(&s.hist[wg_local][bin], (uint32_t) __popc(hm));
    }
}

// One thread per work group publishes to the 12-slice global buffer exactly like
// colorspace.c:1329-1348. (wgs_x, wgs_y) = work groups per CTA tile; (wg0x, wg0y) = global
// index of the CTA's first work group; (nwgx, nwgy) = work groups in the whole dispatch.
PLB_DEV void plb_peak_finish(plb_peak_smem &s, const plb_op_peak &p, int wgs_x, int wgs_y,
                             int wg0x, int wg0y, int nwgx, int nwgy, uint32_t wg_size)
{
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    uint32_t *buf = p.buf;
    uint32_t *wg_count = buf, *wg_active = buf + PLB_PEAK_SLICES,
             *sum_pq = buf + 2 * PLB_PEAK_SLICES, *max_pq = buf + 3 * PLB_PEAK_SLICES,
             *hist = buf + 4 * PLB_PEAK_SLICES;
    const int nwg = wgs_x * wgs_y;
    if (p.use_histogram) {
        for (int i = tid; i < nwg * PLB_HIST_BINS; i += nt) {
            const int wl = i / PLB_HIST_BINS, bin = i - wl * PLB_HIST_BINS;
            const int wx = wg0x + wl % wgs_x, wy = wg0y + wl / wgs_x;
            if (wx >= nwgx || wy >= nwgy)
                continue;
            const uint32_t slice = (uint32_t) (wy * nwgx + wx) % PLB_PEAK_SLICES;
            uint32_t v = s.hist[wl][bin];
            if (bin == 0 && p.cutoff != 0.0f)
                v -= s.black[wl];
            if (v)
                atomicAdd(&hist[slice * PLB_HIST_BINS + bin], v);
        }
    }
    for (int wl = tid; wl < nwg; wl += nt) {
        const int wx = wg0x + wl % wgs_x, wy = wg0y + wl / wgs_x;
        if (wx >= nwgx || wy >= nwgy)
            continue;
        const uint32_t slice = (uint32_t) (wy * nwgx + wx) % PLB_PEAK_SLICES;
        const uint32_t num = wg_size - s.black[wl];
        atomicAdd(&wg_count[slice], 1u);
        atomicAdd(&wg_active[slice], min(num, 1u));
        if (num > 0u) {
            atomicAdd(&sum_pq[slice], s.sum[wl] / num);
            atomicMax(&max_pq[slice], s.mx[wl]);
        }
    }
}
