// Device helpers shared by all kernels: texel load/store with format conversion, sampling,
// pixel coordinates. All kernels are compiled with -fmad=false; every fused multiply-add is
// written explicitly (__fmaf_rn) so that the fp32 operation sequence is exactly the one the
// CPU oracle performs (oracle/pl_oracle.h "arithmetic conventions").
#pragma once

#include "internal.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define PLB_DEV __device__ __forceinline__

PLB_DEV float plb_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
PLB_DEV float plb_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
PLB_DEV int plb_clampi(int x, int lo, int hi) { return min(max(x, lo), hi); }

// ---------------------------------------------------------------------------------------
// Texel access. Conversions: unorm -> v / (2^n - 1) with an IEEE division (what texture units
// and the oracle do), half -> exact widening.
// ---------------------------------------------------------------------------------------

// u / 255.0f and u / 65535.0f as three FP32 operations: q0 = u * fl(1/max), one residual
// correction. Equal to the IEEE division for EVERY 8 / 16-bit code value (proved exhaustively
// in tests/test_unorm_division.py); the division sequence costs about three times as much.
PLB_DEV float plb_unorm8_to_float(unsigned u)
{
    const float uf = (float) u, c = 1.0f / 255.0f;
    const float q0 = uf * c;
    return __fmaf_rn(__fmaf_rn(-q0, 255.0f, uf), c, q0);
}
PLB_DEV float plb_unorm16_to_float(unsigned u)
{
    const float uf = (float) u, c = 1.0f / 65535.0f;
    const float q0 = uf * c;
    return __fmaf_rn(__fmaf_rn(-q0, 65535.0f, uf), c, q0);
}

template <int KIND, int COMPS>
PLB_DEV float4 plb_load_t(const plb_tex_desc &t, int x, int y)
{
    const char *row = (const char *) t.ptr + (size_t) y * t.pitch;
    float v[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    if (KIND == PLB_FK_UNORM8) {
        const uint8_t *p = (const uint8_t *) row + x * COMPS;
        if (COMPS == 4) {
            uchar4 u = *(const uchar4 *) p;
            v[0] = plb_unorm8_to_float(u.x); v[1] = plb_unorm8_to_float(u.y); v[2] = plb_unorm8_to_float(u.z); v[3] = plb_unorm8_to_float(u.w);
        } else {
            #pragma unroll
            for (int c = 0; c < COMPS; c++)
                v[c] = plb_unorm8_to_float(p[c]);
        }
    } else if (KIND == PLB_FK_UNORM16) {
        const uint16_t *p = (const uint16_t *) row + x * COMPS;
        if (COMPS == 4) {
            ushort4 u = *(const ushort4 *) p;
            v[0] = plb_unorm16_to_float(u.x); v[1] = plb_unorm16_to_float(u.y); v[2] = plb_unorm16_to_float(u.z); v[3] = plb_unorm16_to_float(u.w);
        } else if (COMPS == 2) {
            ushort2 u = *(const ushort2 *) p;
            v[0] = plb_unorm16_to_float(u.x); v[1] = plb_unorm16_to_float(u.y);
        } else {
            #pragma unroll
            for (int c = 0; c < COMPS; c++)
                v[c] = plb_unorm16_to_float(p[c]);
        }
    } else if (KIND == PLB_FK_HALF) {
        const __half *p = (const __half *) row + x * COMPS;
        if (COMPS == 4) {
            uint2 raw = *(const uint2 *) p;
            __half2 a = *(__half2 *) &raw.x, b = *(__half2 *) &raw.y;
            float2 fa = __half22float2(a), fb = __half22float2(b);
            v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
        } else if (COMPS == 2) {
            float2 f = __half22float2(*(const __half2 *) p);
            v[0] = f.x; v[1] = f.y;
        } else {
            #pragma unroll
            for (int c = 0; c < COMPS; c++)
                v[c] = __half2float(p[c]);
        }
    } else { // float
        const float *p = (const float *) row + x * COMPS;
        if (COMPS == 4) {
            float4 f = *(const float4 *) p;
            return f;
        } else if (COMPS == 2) {
            float2 f = *(const float2 *) p;
            v[0] = f.x; v[1] = f.y;
        } else {
            #pragma unroll
            for (int c = 0; c < COMPS; c++)
                v[c] = p[c];
        }
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

// Runtime-dispatched fetch (uniform branches). Kept inline: as a real function the call
// overhead cost 10-50 % in the fetch-heavy kernels (deband, plane merges) on B200.
PLB_DEV float4 plb_load1(const plb_tex_desc &t, int x, int y)
{
#define L(K) \
    switch (t.comps) { \
    case 1: return plb_load_t<K, 1>(t, x, y); \
    case 2: return plb_load_t<K, 2>(t, x, y); \
    case 3: return plb_load_t<K, 3>(t, x, y); \
    default: return plb_load_t<K, 4>(t, x, y); \
    }
    switch (t.kind) {
    case PLB_FK_UNORM8:  L(PLB_FK_UNORM8)
    case PLB_FK_UNORM16: L(PLB_FK_UNORM16)
    case PLB_FK_HALF:    L(PLB_FK_HALF)
    default:             L(PLB_FK_FLOAT)
    }
#undef L
}

PLB_DEV float4 plb_load(const plb_tex_desc &t, int x, int y)
{
    if (t.ptr2) { // planar pair view
        plb_tex_desc a = t;
        a.comps = 1;
        const float4 va = plb_load1(a, x, y);
        a.ptr = t.ptr2;
        a.pitch = t.pitch2;
        const float4 vb = plb_load1(a, x, y);
        return make_float4(va.x, vb.x, 0.0f, 1.0f);
    }
    return plb_load1(t, x, y);
}


// Texel index under the sampler's address mode (GL/Vulkan CLAMP_TO_EDGE, REPEAT,
// MIRRORED_REPEAT); the mode is launch-uniform.
PLB_DEV int plb_wrap_index(int i, int n, int mode)
{
    if (mode == 0)
        return plb_clampi(i, 0, n - 1);
    if ((unsigned) i < (unsigned) n)
        return i;
    if (mode == 2 && i >= -n && i < 2 * n) // first reflection: no division
        return i < 0 ? -1 - i : 2 * n - 1 - i;
    const int period = mode == 1 ? n : 2 * n;
    i %= period;
    if (i < 0)
        i += period;
    return i < n ? i : period - 1 - i;
}

PLB_DEV float4 plb_load_clamped(const plb_tex_desc &t, int x, int y)
{
    return plb_load(t, plb_wrap_index(x, t.w, t.addr), plb_wrap_index(y, t.h, t.addr));
}

// textureLod(NEAREST) at normalised coords
PLB_DEV float4 plb_tex_nearest(const plb_tex_desc &t, float u, float v)
{
    int x = (int) floorf(u * (float) t.w), y = (int) floorf(v * (float) t.h);
    return plb_load_clamped(t, x, y);
}

// The same fetch with the texture layout known at compile time. FAST 1: one r16 (unorm) plane,
// clamp addressing; FAST 2: planar pair view of two r16 planes (plb_tex_desc.ptr2); FAST 0: any.
// plb_tex_fast_kind() classifies a descriptor (launch-uniform). Values equal plb_tex_nearest's.
PLB_DEV int plb_tex_fast_kind(const plb_tex_desc &t)
{
    if (t.kind == PLB_FK_HALF && t.comps == 4 && !t.ptr2)
        return 3; // rgba16hf (FBOs); used for 1:1 fetches only, see plb_load_rgba16hf
    if (t.kind != PLB_FK_UNORM16 || t.addr != 0)
        return 0;
    if (t.ptr2)
        return 2;
    return t.comps == 1 ? 1 : 0;
}

// texel (x, y) of an rgba16hf texture, in range (= plb_load_t<PLB_FK_HALF, 4>)
PLB_DEV float4 plb_load_rgba16hf(const plb_tex_desc &t, int x, int y)
{
    return plb_load_t<PLB_FK_HALF, 4>(t, x, y);
}

template <int FAST>
PLB_DEV float4 plb_tex_nearest_t(const plb_tex_desc &t, float u, float v)
{
    if (FAST == 0 || FAST == 3)
        return plb_tex_nearest(t, u, v);
    int x = (int) floorf(u * (float) t.w), y = (int) floorf(v * (float) t.h);
    x = plb_clampi(x, 0, t.w - 1);
    y = plb_clampi(y, 0, t.h - 1);
    const unsigned a = *(const uint16_t *) ((const char *) t.ptr + (size_t) y * t.pitch + 2 * x);
    if (FAST == 1)
        return make_float4(plb_unorm16_to_float(a), 0.0f, 0.0f, 1.0f);
    const unsigned b = *(const uint16_t *) ((const char *) t.ptr2 + (size_t) y * t.pitch2 + 2 * x);
    return make_float4(plb_unorm16_to_float(a), plb_unorm16_to_float(b), 0.0f, 1.0f);
}

PLB_DEV float4 plb_tex_linear(const plb_tex_desc &t, float u, float v)
{
    float tx = u * (float) t.w - 0.5f, ty = v * (float) t.h - 0.5f;
    float bx = floorf(tx), by = floorf(ty);
    float fx = tx - bx, fy = ty - by;
    int x = (int) bx, y = (int) by;
    float4 a = plb_load_clamped(t, x, y), b = plb_load_clamped(t, x + 1, y);
    float4 c = plb_load_clamped(t, x, y + 1), d = plb_load_clamped(t, x + 1, y + 1);
    float4 r;
    r.x = plb_mix(plb_mix(a.x, b.x, fx), plb_mix(c.x, d.x, fx), fy);
    r.y = plb_mix(plb_mix(a.y, b.y, fx), plb_mix(c.y, d.y, fx), fy);
    r.z = plb_mix(plb_mix(a.z, b.z, fx), plb_mix(c.z, d.z, fx), fy);
    r.w = plb_mix(plb_mix(a.w, b.w, fx), plb_mix(c.w, d.w, fx), fy);
    return r;
}

// LINEAR fetch whose coordinate is a texel centre along one axis (the other axis of a separable
// pass that maps 1:1): that axis takes its texel (weights quantised to 1/256 by a texture unit:
// 1, 0), the filtered axis interpolates. axis = 0: interpolate along x; 1: along y.
PLB_DEV float4 plb_tex_linear_axis(const plb_tex_desc &t, float u, float v, int axis)
{
    float4 a, b;
    float f;
    if (axis == 0) {
        const float tx = u * (float) t.w - 0.5f, bx = floorf(tx);
        const int y = (int) floorf(v * (float) t.h);
        f = tx - bx;
        a = plb_load_clamped(t, (int) bx, y);
        b = plb_load_clamped(t, (int) bx + 1, y);
    } else {
        const float ty = v * (float) t.h - 0.5f, by = floorf(ty);
        const int x = (int) floorf(u * (float) t.w);
        f = ty - by;
        a = plb_load_clamped(t, x, (int) by);
        b = plb_load_clamped(t, x, (int) by + 1);
    }
    return make_float4(plb_mix(a.x, b.x, f), plb_mix(a.y, b.y, f), plb_mix(a.z, b.z, f), plb_mix(a.w, b.w, f));
}

// Store one texel. unorm: round-to-nearest-even of clamp(v)*max (GL/Vulkan conversion rule);
// half: RNE narrowing.
PLB_DEV uint32_t plb_unorm(float v, float maxv)
{
    return (uint32_t) rintf(plb_clamp(v, 0.0f, 1.0f) * maxv);
}

PLB_DEV void plb_store(const plb_tex_desc &t, int x, int y, float4 c)
{
    char *row = (char *) t.ptr + (size_t) y * t.pitch;
    const float v[4] = {c.x, c.y, c.z, c.w};
    switch (t.kind) {
    case PLB_FK_UNORM8: {
        uint8_t *p = (uint8_t *) row + x * t.comps;
        if (t.comps == 4) {
            *(uchar4 *) p = make_uchar4(plb_unorm(v[0], 255.0f), plb_unorm(v[1], 255.0f),
                                        plb_unorm(v[2], 255.0f), plb_unorm(v[3], 255.0f));
        } else {
            for (int k = 0; k < t.comps; k++)
                p[k] = plb_unorm(v[k], 255.0f);
        }
        break;
    }
    case PLB_FK_UNORM16: {
        uint16_t *p = (uint16_t *) row + x * t.comps;
        if (t.comps == 4) {
            *(ushort4 *) p = make_ushort4(plb_unorm(v[0], 65535.0f), plb_unorm(v[1], 65535.0f),
                                          plb_unorm(v[2], 65535.0f), plb_unorm(v[3], 65535.0f));
        } else {
            for (int k = 0; k < t.comps; k++)
                p[k] = plb_unorm(v[k], 65535.0f);
        }
        break;
    }
    case PLB_FK_HALF: {
        __half *p = (__half *) row + x * t.comps;
        if (t.comps == 4) {
            __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
            uint2 raw;
            raw.x = *(uint32_t *) &a;
            raw.y = *(uint32_t *) &b;
            *(uint2 *) p = raw;
        } else if (t.comps == 2) {
            *(__half2 *) p = __floats2half2_rn(v[0], v[1]);
        } else {
            for (int k = 0; k < t.comps; k++)
                p[k] = __float2half_rn(v[k]);
        }
        break;
    }
    default: {
        float *p = (float *) row + x * t.comps;
        if (t.comps == 4) {
            *(float4 *) p = c;
        } else {
            for (int k = 0; k < t.comps; k++)
                p[k] = v[k];
        }
        break;
    }
    }
}

// Store four horizontally consecutive RGBA texels (x % 4 == 0 and a 256-byte aligned pitch make
// the 16-byte vector stores legal); falls back to scalar stores otherwise. Same conversions as
// plb_store.
PLB_DEV void plb_store4(const plb_tex_desc &t, int x, int y, const float4 c[4])
{
    char *row = (char *) t.ptr + (size_t) y * t.pitch;
    if (t.comps == 4 && !(x & 3) && t.kind == PLB_FK_HALF) {
        uint4 lo, hi;
        __half2 h;
        h = __floats2half2_rn(c[0].x, c[0].y); lo.x = *(uint32_t *) &h;
        h = __floats2half2_rn(c[0].z, c[0].w); lo.y = *(uint32_t *) &h;
        h = __floats2half2_rn(c[1].x, c[1].y); lo.z = *(uint32_t *) &h;
        h = __floats2half2_rn(c[1].z, c[1].w); lo.w = *(uint32_t *) &h;
        h = __floats2half2_rn(c[2].x, c[2].y); hi.x = *(uint32_t *) &h;
        h = __floats2half2_rn(c[2].z, c[2].w); hi.y = *(uint32_t *) &h;
        h = __floats2half2_rn(c[3].x, c[3].y); hi.z = *(uint32_t *) &h;
        h = __floats2half2_rn(c[3].z, c[3].w); hi.w = *(uint32_t *) &h;
        uint4 *p = (uint4 *) (row + (size_t) x * 8);
        p[0] = lo;
        p[1] = hi;
    } else if (t.comps == 4 && !(x & 3) && t.kind == PLB_FK_UNORM8) {
        uint4 v;
        uint32_t *o = &v.x;
        #pragma unroll
        for (int i = 0; i < 4; i++)
            o[i] = plb_unorm(c[i].x, 255.0f) | (plb_unorm(c[i].y, 255.0f) << 8) |
                   (plb_unorm(c[i].z, 255.0f) << 16) | (plb_unorm(c[i].w, 255.0f) << 24);
        *(uint4 *) (row + (size_t) x * 4) = v;
    } else if (t.comps == 4 && !(x & 3) && t.kind == PLB_FK_UNORM16) {
        uint4 *p = (uint4 *) (row + (size_t) x * 8);
        #pragma unroll
        for (int i = 0; i < 2; i++) {
            uint4 v;
            v.x = plb_unorm(c[2*i].x, 65535.0f) | (plb_unorm(c[2*i].y, 65535.0f) << 16);
            v.y = plb_unorm(c[2*i].z, 65535.0f) | (plb_unorm(c[2*i].w, 65535.0f) << 16);
            v.z = plb_unorm(c[2*i+1].x, 65535.0f) | (plb_unorm(c[2*i+1].y, 65535.0f) << 16);
            v.w = plb_unorm(c[2*i+1].z, 65535.0f) | (plb_unorm(c[2*i+1].w, 65535.0f) << 16);
            p[i] = v;
        }
    } else {
        #pragma unroll
        for (int i = 0; i < 4; i++)
            plb_store(t, x + i, y, c[i]);
    }
}

// ---------------------------------------------------------------------------------------
// Pixel coordinates — same two modes as the oracle (oracle/pl_oracle.c coord_pos/coord_texel);
// LITERAL restates src/dispatch.c:1038-1062 in fp32.
// ---------------------------------------------------------------------------------------

struct plb_coord {
    int mode;
    // EXACT mode, the rect is the whole texture and maps 1:1 onto the output: sample positions
    // are exactly the texel centres, where a bilinear fetch returns the texel itself (texture
    // units quantise the weights to 1/256: 1, 0, 0, 0) — direct and bilinear sampling coincide
    int ident;
    float px0, px1, py0, py1;
    float osx, osy;
    float sizex, sizey;
    float ptx, pty;
    double dx0, dxs, dy0, dys;
};

static inline plb_coord plb_coord_init(int mode, const plb_tex_desc &src, pl_rect2df rc, int dw, int dh)
{
    plb_coord c;
    c.mode = mode;
    float sx = (float) (1.0 / src.w), sy = (float) (1.0 / src.h);
    c.ptx = sx; c.pty = sy;
    c.px0 = sx * rc.x0; c.px1 = sx * rc.x1;
    c.py0 = sy * rc.y0; c.py1 = sy * rc.y1;
    c.osx = (float) (1.0 / dw);
    c.osy = (float) (1.0 / dh);
    c.sizex = (float) src.w;
    c.sizey = (float) src.h;
    c.dx0 = rc.x0; c.dxs = ((double) rc.x1 - rc.x0) / dw;
    c.dy0 = rc.y0; c.dys = ((double) rc.y1 - rc.y0) / dh;
    // (gx + 0.5) / w lands inside texel gx for every w < 2^22: nearest fetches need no division
    c.ident = mode != PLB_COORD_LITERAL && rc.x0 == 0.0f && rc.y0 == 0.0f && rc.x1 == (float) src.w &&
              rc.y1 == (float) src.h && dw == src.w && dh == src.h && src.w < (1 << 22) &&
              src.h < (1 << 22);
    return c;
}

PLB_DEV void plb_coord_pos(const plb_coord &c, int gx, int gy, float &posx, float &posy)
{
    if (c.mode == PLB_COORD_LITERAL) {
        float fx = c.osx * ((float) gx + 0.5f);
        float fy = c.osy * ((float) gy + 0.5f);
        float mx = plb_mix(c.px0, c.px1, fx);
        posx = plb_mix(mx, mx, fy);
        posy = plb_mix(plb_mix(c.py0, c.py0, fx), plb_mix(c.py1, c.py1, fx), fy);
    } else if (c.ident) {
        // (gx + 0.5) / w: the fp32 quotient of two exactly representable operands is the float
        // nearest to the binary64 quotient below (a ratio of integers < 2^24 never comes within
        // 2^-53 of a float midpoint), and costs a quarter of the binary64 division
        posx = __fdiv_rn((float) gx + 0.5f, c.sizex);
        posy = __fdiv_rn((float) gy + 0.5f, c.sizey);
    } else {
        posx = (float) ((c.dx0 + c.dxs * (gx + 0.5)) / c.sizex);
        posy = (float) ((c.dy0 + c.dys * (gy + 0.5)) / c.sizey);
    }
}

PLB_DEV void plb_coord_texel(const plb_coord &c, int gx, int gy, int &jx, int &jy, float &fx, float &fy)
{
    if (c.mode == PLB_COORD_LITERAL) {
        float px, py;
        plb_coord_pos(c, gx, gy, px, py);
        float tx = px * c.sizex - 0.5f, ty = py * c.sizey - 0.5f;
        float bx = floorf(tx), by = floorf(ty);
        jx = (int) bx; jy = (int) by;
        fx = tx - bx; fy = ty - by;
    } else {
        double tx = c.dx0 + c.dxs * (gx + 0.5) - 0.5;
        double ty = c.dy0 + c.dys * (gy + 0.5) - 0.5;
        double bx = floor(tx), by = floor(ty);
        jx = (int) bx; jy = (int) by;
        fx = (float) (tx - bx); fy = (float) (ty - by);
    }
}

// 1-D linear LUT, array form of src/shaders/lut.c:700-715
PLB_DEV float plb_lut1d(const float *lut, int n, float x)
{
    float fpos = plb_clamp(x, 0.0f, 1.0f) * (float) (n - 1);
    float fbase = floorf(fpos), fceil = ceilf(fpos);
    return plb_mix(lut[(int) fbase], lut[(int) fceil], fpos - fbase);
}

// sh_prng (src/shaders.c:965-998), pcg3d
PLB_DEV float3 plb_prng(uint3 &s)
{
    s.x = 1664525u * s.x + 1013904223u;
    s.y = 1664525u * s.y + 1013904223u;
    s.z = 1664525u * s.z + 1013904223u;
    s.x += s.y * s.z;
    s.y += s.z * s.x;
    s.z += s.x * s.y;
    s.x ^= s.x >> 16; s.y ^= s.y >> 16; s.z ^= s.z >> 16;
    s.x += s.y * s.z;
    s.y += s.z * s.x;
    s.z += s.x * s.y;
    const float k = 1.0f / 4294967296.0f;
    return make_float3(__uint2float_rn(s.x) * k, __uint2float_rn(s.y) * k, __uint2float_rn(s.z) * k);
}
