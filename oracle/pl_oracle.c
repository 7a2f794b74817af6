/*
 * TEST INFRASTRUCTURE — see pl_oracle.h for scope, conventions and pinning status.
 * Part 1: coordinates, samplers (polar / ortho / direct), PRNG, debanding, storage rounding.
 *
 * Compile: gcc -std=c11 -O2 -fopenmp -ffp-contract=off -fno-math-errno -fPIC -shared
 */
#define _GNU_SOURCE
#include "pl_oracle.h"
#include "pl_oracle_math.h"

#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ------------------------------------------------------------------------------------------
// GLSL helpers with the rounding conventions documented in pl_oracle.h
// ------------------------------------------------------------------------------------------

static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float fractf(float x) { return x - floorf(x); }
static inline float lengthf2(float x, float y) { return sqrtf(fmaf(x, x, y * y)); }
static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : x > hi ? hi : x; }
static inline float smoothstepf(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// 1-D linear LUT, array form of src/shaders/lut.c:700-715
static inline float lut1d_linear(const float *lut, int n, float x)
{
    float fpos = clampf(x, 0.0f, 1.0f) * (float) (n - 1);
    float fbase = floorf(fpos), fceil = ceilf(fpos);
    float fc = fpos - fbase;
    return mixf(lut[(int) fbase], lut[(int) fceil], fc);
}

// ------------------------------------------------------------------------------------------
// Pixel coordinates. sh_bind (src/shaders.c:513-571) turns the source rect into a vec2
// vertex attribute whose corner values are rect/tex_size; the compute path of
// src/dispatch.c:1038-1062 evaluates it per invocation as
//   frag_map = out_scale * (gid + 0.5);  pos = mix(mix(p0,p1,fx), mix(p2,p3,fx), fy)
// Returns t = pos * size - 0.5 (texel-space position of the sample, minus half), which is
// what every sampler derives its base texel and sub-texel fraction from.
// ------------------------------------------------------------------------------------------

struct coord_ctx {
    int mode;
    float px0, px1, py0, py1; // corner values (literal mode)
    float osx, osy;           // out_scale
    float sizex, sizey;
    float ptx, pty;
    double dx0, dxs, dy0, dys; // exact mode: t = d0 + ds*(gid+0.5) - 0.5
};

static struct coord_ctx coord_init(int mode, const struct orc_tex *src, struct orc_rect rc,
                                   int dw, int dh)
{
    struct coord_ctx c = { .mode = mode };
    float sx = (float) (1.0 / src->w), sy = (float) (1.0 / src->h);
    c.ptx = sx;
    c.pty = sy;
    c.px0 = sx * rc.x0; c.px1 = sx * rc.x1;
    c.py0 = sy * rc.y0; c.py1 = sy * rc.y1;
    c.osx = (float) (1.0 / dw);
    c.osy = (float) (1.0 / dh);
    c.sizex = (float) src->w;
    c.sizey = (float) src->h;
    c.dx0 = rc.x0; c.dxs = ((double) rc.x1 - rc.x0) / dw;
    c.dy0 = rc.y0; c.dys = ((double) rc.y1 - rc.y0) / dh;
    return c;
}

// normalised position (pos) of invocation (gx, gy)
static inline void coord_pos(const struct coord_ctx *c, int gx, int gy, float *posx, float *posy)
{
    if (c->mode == ORC_COORD_LITERAL) {
        float fx = c->osx * ((float) gx + 0.5f);
        float fy = c->osy * ((float) gy + 0.5f);
        float mx = mixf(c->px0, c->px1, fx);
        float my0 = c->py0, my1 = c->py1;
        *posx = mixf(mx, mx, fy);
        *posy = mixf(mixf(my0, my0, fx), mixf(my1, my1, fx), fy);
    } else {
        *posx = (float) ((c->dx0 + c->dxs * (gx + 0.5)) / c->sizex);
        *posy = (float) ((c->dy0 + c->dys * (gy + 0.5)) / c->sizey);
    }
}

// t = pos*size - 0.5 split into integer base texel and fraction
static inline void coord_texel(const struct coord_ctx *c, int gx, int gy,
                               int *jx, int *jy, float *fx, float *fy)
{
    if (c->mode == ORC_COORD_LITERAL) {
        float px, py;
        coord_pos(c, gx, gy, &px, &py);
        float tx = px * c->sizex - 0.5f, ty = py * c->sizey - 0.5f;
        float bx = floorf(tx), by = floorf(ty);
        *jx = (int) bx; *jy = (int) by;
        *fx = tx - bx;  *fy = ty - by;
    } else {
        double tx = c->dx0 + c->dxs * (gx + 0.5) - 0.5;
        double ty = c->dy0 + c->dys * (gy + 0.5) - 0.5;
        double bx = floor(tx), by = floor(ty);
        *jx = (int) bx; *jy = (int) by;
        *fx = (float) (tx - bx); *fy = (float) (ty - by);
    }
}

// texel index under the sampler's address mode (CLAMP_TO_EDGE / REPEAT / MIRRORED_REPEAT)
static inline int wrap_index(int i, int n, int mode)
{
    if (mode == 0)
        return clampi(i, 0, n - 1);
    const int period = mode == 1 ? n : 2 * n;
    i %= period;
    if (i < 0)
        i += period;
    return i < n ? i : period - 1 - i;
}

static inline const float *texel(const struct orc_tex *t, int x, int y)
{
    x = wrap_index(x, t->w, t->addr);
    y = wrap_index(y, t->h, t->addr);
    return t->data + ((size_t) y * t->w + x) * t->comps;
}

// textureLod with NEAREST at normalised coordinate (u, v): floor(u*size), clamp-to-edge
static inline const float *tex_nearest(const struct orc_tex *t, float u, float v)
{
    int x = (int) floorf(u * (float) t->w), y = (int) floorf(v * (float) t->h);
    return texel(t, x, y);
}

// textureLod with LINEAR
static inline void tex_linear(const struct orc_tex *t, float u, float v, float out[4])
{
    float tx = u * (float) t->w - 0.5f, ty = v * (float) t->h - 0.5f;
    float bx = floorf(tx), by = floorf(ty);
    float fx = tx - bx, fy = ty - by;
    int x = (int) bx, y = (int) by;
    const float *a = texel(t, x, y), *b = texel(t, x + 1, y);
    const float *c = texel(t, x, y + 1), *d = texel(t, x + 1, y + 1);
    for (int k = 0; k < t->comps; k++)
        out[k] = mixf(mixf(a[k], b[k], fx), mixf(c[k], d[k], fx), fy);
}

// LINEAR fetch at a texel centre along one axis: that axis takes its texel, the other one
// interpolates (axis 0: along x, 1: along y)
static inline void tex_linear_axis(const struct orc_tex *t, float u, float v, int axis, float out[4])
{
    const float *a, *b;
    float f;
    if (axis == 0) {
        float tx = u * (float) t->w - 0.5f, bx = floorf(tx);
        int y = (int) floorf(v * (float) t->h);
        f = tx - bx;
        a = texel(t, (int) bx, y);
        b = texel(t, (int) bx + 1, y);
    } else {
        float ty = v * (float) t->h - 0.5f, by = floorf(ty);
        int x = (int) floorf(u * (float) t->w);
        f = ty - by;
        a = texel(t, x, (int) by);
        b = texel(t, x, (int) by + 1);
    }
    for (int k = 0; k < t->comps; k++)
        out[k] = mixf(a[k], b[k], f);
}

float orc_tex_linear_r(const struct orc_tex *t, float u, float v)
{
    float out[4] = {0, 0, 0, 1};
    tex_linear(t, u, v, out);
    return out[0];
}

// ------------------------------------------------------------------------------------------
// pl_shader_deinterlace — src/shaders/deinterlacing.c:32-370 (weave / bob / yadif / bwdif on
// interleaved frames). GET(TEX, X, Y) = NEAREST fetch with MIRROR addressing at pos + pt*(X, Y),
// i.e. texel (x + X, y + Y) reflected at the borders. out: vec4(0,0,0,1) with the masked
// components replaced. GLSL statement by statement, fp32, no contraction.
// ------------------------------------------------------------------------------------------

static inline float dget(const struct orc_tex *t, int x, int y, int k)
{
    x = wrap_index(x, t->w, 2);
    y = wrap_index(y, t->h, 2);
    return k < t->comps ? t->data[((size_t) y * t->w + x) * t->comps + k] : (k == 3 ? 1.0f : 0.0f);
}

static float yadif_spatial(const float up[7], const float dn[7])
{
    const float a = up[0], b = up[1], c = up[2], d = up[3], e = up[4], f = up[5], g = up[6];
    const float h = dn[0], i = dn[1], j = dn[2], k = dn[3], l = dn[4], m = dn[5], n = dn[6];
    float spatial_pred = (d + k) / 2.0f;
    float spatial_score = fabsf(c - j) + fabsf(d - k) + fabsf(e - l) - 0.003922f; // "%f" of 1/255
    float score = fabsf(b - k) + fabsf(c - l) + fabsf(d - m);
    if (score < spatial_score) {
        spatial_pred = (c + l) / 2.0f;
        spatial_score = score;
        score = fabsf(a - l) + fabsf(b - m) + fabsf(c - n);
        if (score < spatial_score) {
            spatial_pred = (b + m) / 2.0f;
            spatial_score = score;
        }
    }
    score = fabsf(d - i) + fabsf(e - j) + fabsf(f - k);
    if (score < spatial_score) {
        spatial_pred = (e + j) / 2.0f;
        spatial_score = score;
        score = fabsf(e - h) + fabsf(f - i) + fabsf(g - j);
        if (score < spatial_score) {
            spatial_pred = (f + i) / 2.0f;
            spatial_score = score;
        }
    }
    return spatial_pred;
}

void orc_deinterlace(const struct orc_tex *cur, const struct orc_tex *prev_in,
                     const struct orc_tex *next_in, int field, int first_field, int algo,
                     int skip_spatial_check, unsigned comp_mask, float *dst)
{
    const int w = cur->w, h = cur->h;
    if (!first_field)
        first_field = 1;
    int intra_only = algo != 2;
    if (algo == 3)
        intra_only = (!prev_in && field == first_field) || (!next_in && field != first_field);
    const struct orc_tex *prev = (!intra_only && prev_in) ? prev_in : cur;
    const struct orc_tex *next = (!intra_only && next_in) ? next_in : cur;
    const struct orc_tex *prev2 = field == first_field ? prev : cur;
    const struct orc_tex *next2 = field == first_field ? cur : next;
    comp_mask = (comp_mask ? comp_mask : 0xFu) & ((1u << cur->comps) - 1u);

    #pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float res[4];
            const bool keep = field == 0 || (y % 2) == (field == 1 ? 0 : 1) || algo == 0;
            for (int k = 0; k < 4; k++) {
                if (keep) {
                    res[k] = dget(cur, x, y, k);
                } else if (algo == 1) {
                    res[k] = dget(cur, x, y + (field == 1 ? -1 : 1), k);
                } else if (algo == 2) {
                    float up[7], dn[7];
                    for (int i = 0; i < 7; i++) {
                        up[i] = dget(cur, x + i - 3, y - 1, k);
                        dn[i] = dget(cur, x + i - 3, y + 1, k);
                    }
                    float sp = yadif_spatial(up, dn);
                    const float A = dget(prev, x, y - 1, k), B = dget(prev, x, y + 1, k);
                    const float C = dget(prev2, x, y - 2, k), D = dget(prev2, x, y, k), E = dget(prev2, x, y + 2, k);
                    const float F = up[3], G = dn[3];
                    const float H = dget(next2, x, y - 2, k), I = dget(next2, x, y, k), J = dget(next2, x, y + 2, k);
                    const float K = dget(next, x, y - 1, k), L = dget(next, x, y + 1, k);
                    const float p0 = (C + H) / 2.0f, p1 = F, p2 = (D + I) / 2.0f, p3 = G, p4 = (E + J) / 2.0f;
                    const float tdiff0 = fabsf(D - I) / 2.0f;
                    const float tdiff1 = (fabsf(A - F) + fabsf(B - G)) / 2.0f;
                    const float tdiff2 = (fabsf(K - F) + fabsf(G - L)) / 2.0f;
                    float diff = fmaxf(tdiff0, fmaxf(tdiff1, tdiff2));
                    if (!skip_spatial_check) {
                        const float maxi = fmaxf(p2 - fminf(p3, p1), fminf(p0 - p1, p4 - p3));
                        const float mini = fminf(p2 - fmaxf(p3, p1), fmaxf(p0 - p1, p4 - p3));
                        diff = fmaxf(diff, fmaxf(mini, -maxi));
                    }
                    if (sp > p2 + diff)
                        sp = p2 + diff;
                    if (sp < p2 - diff)
                        sp = p2 - diff;
                    res[k] = sp;
                } else {
                    const float sp0 = 5077.0f / 8192.0f, sp1 = 981.0f / 8192.0f;
                    const float cu[4] = { dget(cur, x, y - 3, k), dget(cur, x, y - 1, k),
                                          dget(cur, x, y + 1, k), dget(cur, x, y + 3, k) };
                    if (intra_only) {
                        res[k] = sp0 * (cu[1] + cu[2]) - sp1 * (cu[0] + cu[3]);
                        continue;
                    }
                    const float lf0 = 4309.0f / 8192.0f, lf1 = 213.0f / 8192.0f;
                    const float hf0 = 5570.0f / 8192.0f, hf1 = 3801.0f / 8192.0f, hf2 = 1016.0f / 8192.0f;
                    const float pv[2] = { dget(prev, x, y - 1, k), dget(prev, x, y + 1, k) };
                    const float nx[2] = { dget(next, x, y - 1, k), dget(next, x, y + 1, k) };
                    float p2v[5], n2v[5];
                    for (int i = 0; i < 5; i++) {
                        p2v[i] = dget(prev2, x, y + 2 * i - 4, k);
                        n2v[i] = dget(next2, x, y + 2 * i - 4, k);
                    }
                    const float s = p2v[2] + n2v[2];
                    const float d = s / 2.0f;
                    const float c = cu[1], e = cu[2];
                    const float tdiff0 = fabsf(p2v[2] - n2v[2]);
                    const float tdiff1 = fabsf(pv[0] - c) + fabsf(pv[1] - e);
                    const float tdiff2 = fabsf(nx[0] - c) + fabsf(nx[1] - e);
                    float diff = fmaxf(tdiff0, fmaxf(tdiff1, tdiff2)) / 2.0f;
                    const bool diff_mask = diff == 0.0f;
                    const float bs = p2v[1] + n2v[1], fs = p2v[3] + n2v[3];
                    const float b = (bs / 2.0f) - c, f = (fs / 2.0f) - c;
                    const float dc = d - c, de = d - e;
                    const float mmax = fmaxf(de, fmaxf(dc, fminf(b, f)));
                    const float mmin = fminf(de, fminf(dc, fmaxf(b, f)));
                    diff = fmaxf(diff, fmaxf(mmin, -mmax));
                    const float single = sp0 * (c + e) - sp1 * (cu[0] + cu[3]);
                    float all = (hf0 * s - hf1 * (bs + fs) + hf2 * (p2v[0] + n2v[0] + p2v[4] + n2v[4])) / 4.0f;
                    all += lf0 * (c + e) - lf1 * (cu[0] + cu[3]);
                    float interpol = fabsf(c - e) > tdiff0 ? all : single;
                    interpol = fminf(fmaxf(interpol, d - diff), d + diff);
                    res[k] = diff_mask ? d : interpol;
                }
            }
            float *o = dst + ((size_t) y * w + x) * 4;
            o[0] = o[1] = o[2] = 0.0f;
            o[3] = 1.0f;
            for (int k = 0; k < 4; k++) {
                if (comp_mask & (1u << k))
                    o[k] = res[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_sample_polar — src/shaders/sampling.c:587-912; per-tap body polar_sample :503-558
// ------------------------------------------------------------------------------------------

// Evaluation extent. A compute dispatch runs whole work groups, so invocations beyond the
// image still execute the sampling code (with positions past the edge and clamped texel
// fetches); only their store is masked (dispatch.c:1068-1075). Peak detection counts those
// invocations. With a non-zero extent the samplers below evaluate ew x eh invocations instead
// of dw x dh (dst then holds ew*eh pixels) while every coordinate still derives from dw x dh.
static int g_ext_w, g_ext_h;
void orc_set_eval_extent(int ew, int eh)
{
    g_ext_w = ew;
    g_ext_h = eh;
}
#define EXT_W(dw) ((g_ext_w > (dw)) ? g_ext_w : (dw))
#define EXT_H(dh) ((g_ext_h > (dh)) ? g_ext_h : (dh))

void orc_sample_polar(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                      const struct orc_polar_params *p, float *dst)
{
    const struct coord_ctx cc = coord_init(p->coord_mode, src, rect, dw, dh);
    const int bound = (int) ceilf(p->radius);        // sampling.c:661
    const float radius = p->radius;
    const bool use_ar_global = p->antiring > 0.0f;
    const float ar_radius = p->radius_zero;

    // Compile-time tap list (sampling.c:776-783 with the culling of :508-522)
    int ntaps = 0;
    int tap_x[256], tap_y[256];
    unsigned char tap_guard[256], tap_ar[256];
    for (int y = 1 - bound; y <= bound; y++) {
        for (int x = 1 - bound; x <= bound; x++) {
            int yy = y > 0 ? y - 1 : y;
            int xx = x > 0 ? x - 1 : x;
            float dmin = sqrt(xx * xx + yy * yy);
            if (dmin >= radius)
                continue;
            tap_x[ntaps] = x;
            tap_y[ntaps] = y;
            tap_guard[ntaps] = dmin >= radius - M_SQRT2;
            tap_ar[ntaps] = use_ar_global && dmin < ar_radius;
            ntaps++;
        }
    }

    const int ew = EXT_W(dw), eh = EXT_H(dh);
    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < eh; gy++) {
        for (int gx = 0; gx < ew; gx++) {
            int jx, jy;
            float fx, fy;
            coord_texel(&cc, gx, gy, &jx, &jy, &fx, &fy);
            float color[4] = {0, 0, 0, 0};
            float wsum = 0.0f;
            float ar[4][2] = {{0}}, wwsum[4][2] = {{0}};
            for (int t = 0; t < ntaps; t++) {
                float ox = (float) tap_x[t], oy = (float) tap_y[t];
                float d = lengthf2(ox - fx, oy - fy);
                if (tap_guard[t] && !(d < radius))
                    continue;
                float w = lut1d_linear(p->lut, p->lut_entries, d * 1.0f / radius); // sampling.c:529
                wsum += w;
                const float *c = texel(src, jx + tap_x[t], jy + tap_y[t]);
                for (int k = 0; k < 4; k++) {
                    if (!(p->comp_mask & (1 << k)) || k >= src->comps)
                        continue;
                    color[k] = fmaf(w, c[k], color[k]);
                    if (tap_ar[t] && d <= ar_radius) {
                        // anti-ringing accumulators, sampling.c:539-555
                        float cc0 = 1.0f - p->scale * c[k], cc1 = p->scale * c[k];
                        float ww0 = cc0 + 0.10f, ww1 = cc1 + 0.10f;
                        for (int i = 0; i < 5; i++) {
                            ww0 = ww0 * ww0;
                            ww1 = ww1 * ww1;
                        }
                        ww0 = w * ww0;
                        ww1 = w * ww1;
                        ar[k][0] = fmaf(ww0, cc0, ar[k][0]);
                        ar[k][1] = fmaf(ww1, cc1, ar[k][1]);
                        wwsum[k][0] += ww0;
                        wwsum[k][1] += ww1;
                    }
                }
            }

            // sampling.c:896-908
            float norm = p->scale / wsum;
            float *o = dst + ((size_t) gy * ew + gx) * 4;
            for (int k = 0; k < 4; k++) {
                bool have = (p->comp_mask & (1 << k)) && k < src->comps;
                float v = norm * color[k];
                if (have && use_ar_global) {
                    float wx = ar[k][0] / wwsum[k][0], wy = ar[k][1] / wwsum[k][1];
                    wx = 1.0f - wx;
                    float w = clampf(v, wx, wy);
                    if (wx > wy)
                        w = wx * 0.5f + wy * 0.5f; // dot(ww, vec2(0.5))
                    v = mixf(v, w, p->antiring);
                }
                o[k] = have ? v : 0.0f;
            }
            if (!(p->comp_mask & 8))
                o[3] = 1.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_sample_ortho2 — sampling.c:950-1104
// The weights texture is rgba32f, width = row_stride/4, height = 256, sampled LINEARly at
// (float(n/4)/denom, fcoord): x hits texel centres exactly, y interpolates between phases.
// ------------------------------------------------------------------------------------------

static inline void ortho_weights(const struct orc_ortho_params *p, int quad, float fcoord,
                                 float ws[4])
{
    // LUT position macro of lut.c:285-295 along y: texel-space = fcoord * (n-1)
    int width = p->row_stride / 4;
    (void) width;
    float fpos = clampf(fcoord, 0.0f, 1.0f) * (float) (p->lut_entries - 1);
    float fbase = floorf(fpos), fceil = ceilf(fpos);
    float fc = fpos - fbase;
    const float *r0 = p->lut + (size_t) fbase * p->row_stride + 4 * quad;
    const float *r1 = p->lut + (size_t) fceil * p->row_stride + 4 * quad;
    for (int k = 0; k < 4; k++)
        ws[k] = mixf(r0[k], r1[k], fc);
}

void orc_sample_ortho(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                      const struct orc_ortho_params *p, float *dst)
{
    const struct coord_ctx cc = coord_init(p->coord_mode, src, rect, dw, dh);
    const int N = p->row_size;
    const int dirx = p->pass == 1, diry = p->pass == 0;
    const bool use_ar = p->antiring > 0 && p->ratio > 1.0f && !p->use_linear;
    const bool aligned = p->coord_mode != ORC_COORD_LITERAL &&
        (p->pass == 0 ? (rect.x0 == 0.0f && rect.x1 == (float) src->w && dw == src->w)
                      : (rect.y0 == 0.0f && rect.y1 == (float) src->h && dh == src->h));

    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < dh; gy++) {
        for (int gx = 0; gx < dw; gx++) {
            float posx, posy;
            coord_pos(&cc, gx, gy, &posx, &posy);
            // fcoord2 = fract(pos*size - 0.5); fcoord = dot(fcoord2, dir)
            float f2x = fractf(posx * cc.sizex - 0.5f), f2y = fractf(posy * cc.sizey - 0.5f);
            float fcoord = f2x * (float) dirx + f2y * (float) diry;
            float ptx = cc.ptx * (float) dirx, pty = cc.pty * (float) diry;
            float off0 = (float) (N / 2 - 1);
            float basex = posx - fcoord * ptx - ptx * off0;
            float basey = posy - fcoord * pty - pty * off0;
            float ca[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0}, lo[4] = {1e9f, 1e9f, 1e9f, 1e9f};
            float ws[4] = {0};
            for (int n = 0; n < N; n += p->use_linear ? 2 : 1) {
                if (n % 4 == 0)
                    ortho_weights(p, n / 4, fcoord, ws);
                float off = (float) n;
                float c[4] = {0, 0, 0, 0};
                float u, v;
                if (p->use_linear) {
                    off += ws[n % 4 + 1];
                    u = basex + ptx * off;
                    v = basey + pty * off;
                    if (aligned) // texel centre along the unfiltered axis: interpolate along dir
                        tex_linear_axis(src, u, v, p->pass == 1 ? 0 : 1, c);
                    else
                        tex_linear(src, u, v, c);
                } else {
                    u = basex + ptx * off;
                    v = basey + pty * off;
                    // setup_src(LINEAR): the texture is bound with linear filtering, but the
                    // coordinate sits on a texel centre along `dir`; when the other axis maps
                    // 1:1 onto the whole texture it is a texel centre there too, and a LINEAR
                    // fetch at a texel centre is the texel (weights quantised to 1/256)
                    if (aligned) {
                        const float *px = tex_nearest(src, u, v);
                        for (int k = 0; k < src->comps; k++)
                            c[k] = px[k];
                    } else {
                        tex_linear(src, u, v, c);
                    }
                }
                if (use_ar && (n == N / 2 - 1 || n == N / 2)) {
                    for (int k = 0; k < src->comps; k++) {
                        lo[k] = fminf(lo[k], c[k]);
                        hi[k] = fmaxf(hi[k], c[k]);
                    }
                }
                for (int k = 0; k < src->comps; k++)
                    ca[k] = fmaf(ws[n % 4], c[k], ca[k]);
            }
            float *o = dst + ((size_t) gy * dw + gx) * 4;
            o[0] = o[1] = o[2] = 0.0f;
            o[3] = 1.0f;
            for (int k = 0; k < 4; k++) {
                if (!(p->comp_mask & (1 << k)) || k >= src->comps)
                    continue;
                float v = ca[k];
                if (use_ar)
                    v = mixf(v, clampf(v, lo[k], hi[k]), p->antiring);
                o[k] = p->scale * v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pl_shader_sample_direct / nearest / bilinear — sampling.c:277-340:
//   color = scale * textureLod(src_tex, pos, 0.0)
// ------------------------------------------------------------------------------------------

void orc_sample_direct(const struct orc_tex *src, struct orc_rect rect, int dw, int dh,
                       int linear, float scale, int coord_mode, float *dst)
{
    const struct coord_ctx cc = coord_init(coord_mode, src, rect, dw, dh);
    const int ew = EXT_W(dw), eh = EXT_H(dh);
    // 1:1 mapping of the whole texture: the sample positions are the texel centres, where a
    // texture unit's bilinear filter (weights quantised to 1/256) returns the texel itself
    if (coord_mode != ORC_COORD_LITERAL && rect.x0 == 0 && rect.y0 == 0 && rect.x1 == src->w &&
        rect.y1 == src->h && dw == src->w && dh == src->h)
        linear = 0;
    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < eh; gy++) {
        for (int gx = 0; gx < ew; gx++) {
            float posx, posy;
            coord_pos(&cc, gx, gy, &posx, &posy);
            // GLSL texture() returns (0,0,0,1) for missing components
            float c[4] = {0, 0, 0, 1};
            if (linear) {
                tex_linear(src, posx, posy, c);
            } else {
                const float *t = tex_nearest(src, posx, posy);
                for (int k = 0; k < src->comps; k++)
                    c[k] = t[k];
            }
            float *o = dst + ((size_t) gy * ew + gx) * 4;
            for (int k = 0; k < 4; k++)
                o[k] = scale * c[k];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Built-in fast scalers — src/shaders/sampling.c:319-478, the GLSL restated statement by
// statement. kind: 0 bicubic (:319-366), 1 hermite (:368-393), 2 gaussian (:395-440),
// 3 oversample (:442-478; rx, ry = scaling ratios, threshold already clamped to [0, 0.5]).
// `pt` is the uniform the host passes: (float) (1.0 / size). textureLod = tex_linear.
// ------------------------------------------------------------------------------------------

static inline float fractf_(float x) { return x - floorf(x); }
static inline float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

void orc_sample_fast(const struct orc_tex *src, struct orc_rect rect, int dw, int dh, int kind,
                     float scale, float rx, float ry, float threshold, int coord_mode, float *dst)
{
    const struct coord_ctx cc = coord_init(coord_mode, src, rect, dw, dh);
    const int ew = EXT_W(dw), eh = EXT_H(dh);
    const float size[2] = { (float) src->w, (float) src->h };
    const float pt[2] = { (float) (1.0 / src->w), (float) (1.0 / src->h) };
    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < eh; gy++) {
        for (int gx = 0; gx < ew; gx++) {
            float pos[2];
            coord_pos(&cc, gx, gy, &pos[0], &pos[1]);
            float color[4] = {0, 0, 0, 1};
            if (kind == 1) {
                float p[2];
                for (int a = 0; a < 2; a++) {
                    const float frac = fractf_(pos[a] * size[a] + 0.5f);
                    const float t = clamp01(frac); // smoothstep(0.0, 1.0, frac)
                    const float sm = t * t * (3.0f - 2.0f * t);
                    p[a] = pos[a] + pt[a] * (sm - frac);
                }
                tex_linear(src, p[0], p[1], color);
            } else if (kind == 3) {
                float p[2];
                const float ratio[2] = { rx, ry };
                for (int a = 0; a < 2; a++) {
                    const float fcoord = fractf_(pos[a] * size[a] - 0.5f);
                    float coeff = (fcoord - 0.5f) * ratio[a];
                    coeff = clamp01(coeff + 0.5f);
                    if (threshold > 0.0f) {
                        if (coeff < threshold)
                            coeff = 0.0f;
                        if (coeff > 1.0f - threshold)
                            coeff = 1.0f;
                    }
                    p[a] = pos[a] + (coeff - fcoord) * pt[a];
                }
                tex_linear(src, p[0], p[1], color);
            } else {
                float g[4], h[4], off[2] = {0, 0}; // vec4 order: (x, y, z, w) = (x01, y01, x23, y23)
                for (int a = 0; a < 2; a++) {
                    if (kind == 0) {
                        const float frac = fractf_(pos[a] * size[a] + 0.5f);
                        const float frac2 = frac * frac, inv = 1.0f - frac, inv2 = inv * inv;
                        const float w0 = (1.0f / 6.0f) * inv2 * inv;
                        const float w1 = (2.0f / 3.0f) - 0.5f * frac2 * (2.0f - frac);
                        const float w2 = (2.0f / 3.0f) - 0.5f * inv2 * (2.0f - inv);
                        const float w3 = (1.0f / 6.0f) * frac2 * frac;
                        g[a] = w0 + w1;
                        g[2 + a] = w2 + w3;
                        h[a] = w1 / g[a] + inv;
                        h[2 + a] = w3 / g[2 + a] + inv;
                        h[a] -= 2.0f;
                    } else {
                        off[a] = -fractf_(pos[a] * size[a] + 0.5f);
                        const float o = off[a], off2 = -2.0f * o * o;
                        const float w0 = pm_expf(off2 + 4.0f * o - 2.0f);
                        const float w1 = pm_expf(off2);
                        const float w2 = pm_expf(off2 - 4.0f * o - 2.0f);
                        const float w3 = pm_expf(off2 - 8.0f * o - 8.0f);
                        g[a] = w0 + w1;
                        g[2 + a] = w2 + w3;
                        h[a] = w1 / g[a];
                        h[2 + a] = w3 / g[2 + a];
                        h[a] -= 1.0f;
                        h[2 + a] += 1.0f;
                    }
                }
                if (kind == 2) {
                    g[0] /= g[0] + g[2];
                    g[1] /= g[1] + g[3];
                    for (int k = 0; k < 4; k++)
                        h[k] = h[k] + off[k & 1];
                }
                const float px = pos[0] + pt[0] * h[0], py = pos[1] + pt[1] * h[1];
                const float pz = pos[0] + pt[0] * h[2], pw = pos[1] + pt[1] * h[3];
                float c00[4] = {0, 0, 0, 1}, c01[4] = {0, 0, 0, 1}, c10[4] = {0, 0, 0, 1}, c11[4] = {0, 0, 0, 1};
                tex_linear(src, px, py, c00);
                tex_linear(src, px, pw, c01);
                tex_linear(src, pz, py, c10);
                tex_linear(src, pz, pw, c11);
                for (int k = 0; k < 4; k++) {
                    const float c0 = mixf(c01[k], c00[k], g[1]);
                    const float c1 = mixf(c11[k], c10[k], g[1]);
                    color[k] = mixf(c1, c0, g[0]);
                }
            }
            float *o = dst + ((size_t) gy * ew + gx) * 4;
            for (int k = 0; k < 4; k++)
                o[k] = scale * color[k];
        }
    }
}

// ------------------------------------------------------------------------------------------
// sh_prng — src/shaders.c:965-998 (pcg3d). float(0xFFFFFFFFu) rounds to 2^32.
// ------------------------------------------------------------------------------------------

void orc_prng(uint32_t s[3], float out[3])
{
    for (int i = 0; i < 3; i++)
        s[i] = 1664525u * s[i] + 1013904223u;
    s[0] += s[1] * s[2];
    s[1] += s[2] * s[0];
    s[2] += s[0] * s[1];
    for (int i = 0; i < 3; i++)
        s[i] ^= s[i] >> 16;
    s[0] += s[1] * s[2];
    s[1] += s[2] * s[0];
    s[2] += s[0] * s[1];
    const float k = 1.0f / 4294967296.0f;
    for (int i = 0; i < 3; i++)
        out[i] = (float) s[i] * k;
}

// ------------------------------------------------------------------------------------------
// pl_shader_deband — sampling.c:183-275. The source is bound NEAREST (setup_src(..., NEAREST)),
// pos is the texel centre of the output pixel (1:1 pass), taps are at pos + pt*(±d.x, ±d.y).
// ------------------------------------------------------------------------------------------

void orc_deband(const struct orc_tex *src, const struct orc_deband_params *p, float *dst)
{
    const int w = src->w, h = src->h;
    const struct orc_rect full = { 0, 0, (float) w, (float) h };
    const struct coord_ctx cc = coord_init(ORC_COORD_EXACT, src, full, w, h);
    uint8_t mask = p->comp_mask & ~0x8u;
    const float radius = p->radius;
    const float threshold = p->threshold / (1000 * p->scale); // sampling.c:225-226
    const float two_pi = (float) (M_PI * 2);                  // printed with %f: 6.283185
    // "%f" prints 6 decimals; the GLSL literal is 6.283185
    const float two_pi_lit = 6.283185f;
    (void) two_pi;
    const float grain_str = (float) (p->grain / (1000.0 * p->scale));

    const int ew = EXT_W(w), eh = EXT_H(h);
    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < eh; gy++) {
        for (int gx = 0; gx < ew; gx++) {
            float posx, posy;
            coord_pos(&cc, gx, gy, &posx, &posy);
            float color[4] = {0, 0, 0, 1};
            const float *c0 = tex_nearest(src, posx, posy);
            for (int k = 0; k < src->comps; k++)
                color[k] = c0[k];
            float *o = dst + ((size_t) gy * ew + gx) * 4;
            if (!mask) {
                for (int k = 0; k < 4; k++)
                    o[k] = color[k] * p->scale;
                continue;
            }

            // uvec3(gl_FragCoord.xy, index): gl_FragCoord = gid + 0.5, truncated
            uint32_t state[3] = { (uint32_t) gx, (uint32_t) gy, p->frame_index };
            float res[4], avg[4];
            for (int k = 0; k < 4; k++)
                res[k] = color[k];

            for (int i = 1; i <= p->iterations; i++) {
                float rnd[3];
                orc_prng(state, rnd);
                // d = RAND.xy * vec2(i * radius, 2pi); d = d.x * vec2(cos(d.y), sin(d.y))
                float dist = rnd[0] * ((float) i * radius);
                float dir = rnd[1] * two_pi_lit;
                float sn, cs;
                pf_sincosf(dir, &sn, &cs);
                float dx = dist * cs, dy = dist * sn;
                const float sgx[4] = {+1, -1, -1, +1}, sgy[4] = {+1, +1, -1, -1};
                for (int k = 0; k < 4; k++)
                    avg[k] = 0.0f;
                for (int t = 0; t < 4; t++) {
                    const float *c = tex_nearest(src, posx + cc.ptx * (sgx[t] * dx),
                                                      posy + cc.pty * (sgy[t] * dy));
                    for (int k = 0; k < src->comps; k++)
                        avg[k] += c[k];
                }
                float bound = threshold / (float) i;
                for (int k = 0; k < 3; k++) {
                    if (!(mask & (1 << k)))
                        continue;
                    float a = avg[k] * 0.25f;
                    float diff = fabsf(res[k] - a);
                    res[k] = diff > bound ? res[k] : a; // mix(avg, res, greaterThan(diff, bound))
                }
            }

            if (p->grain > 0) {
                float rnd[3];
                orc_prng(state, rnd);
                int idx = 0;
                for (int k = 0; k < 3; k++) {
                    if (!(mask & (1 << k)))
                        continue;
                    float bound = p->grain_neutral[idx] / p->scale;
                    float strength = fminf(fabsf(res[k] - bound), grain_str);
                    res[k] = fmaf(strength, rnd[idx] - 0.5f, res[k]);
                    idx++;
                }
            }

            for (int k = 0; k < 3; k++) {
                if (mask & (1 << k))
                    color[k] = res[k];
            }
            for (int k = 0; k < 4; k++)
                o[k] = color[k] * p->scale;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Storage rounding
// ------------------------------------------------------------------------------------------

static inline float round_half(float v) { return (float) (_Float16) v; }

void orc_store_round(float *data, size_t n, int kind, int bits)
{
    if (kind == 2)
        return;
    const float maxv = (float) ((1u << bits) - 1);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        if (kind == 0) {
            data[i] = round_half(data[i]);
        } else {
            float c = clampf(data[i], 0.0f, 1.0f);
            data[i] = rintf(c * maxv) / maxv;
        }
    }
}

void orc_store_unorm(const float *data, size_t n, int bits, uint16_t *out)
{
    const float maxv = (float) ((1u << bits) - 1);
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float c = clampf(data[i], 0.0f, 1.0f); // NaN -> 0 via fmaxf/fminf ordering
        out[i] = (uint16_t) rintf(c * maxv);
    }
}

void orc_store_half(const float *data, size_t n, uint16_t *out)
{
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        _Float16 hval = (_Float16) data[i];
        memcpy(&out[i], &hval, 2);
    }
}

// ------------------------------------------------------------------------------------------
// Double-precision ideal EWA resampler for config C1: weights straight from
// pl_filter_sample (src/filters.c:82-124), no LUT, no cutoff other than `radius`.
// ------------------------------------------------------------------------------------------

void orc_polar_ideal(const struct orc_tex *src, int dw, int dh, double radius,
                     orc_weight_fn fn, void *priv, double *dst)
{
    const int bound = (int) ceil(radius);
    for (int gy = 0; gy < dh; gy++) {
        for (int gx = 0; gx < dw; gx++) {
            double tx = (gx + 0.5) * src->w / dw - 0.5, ty = (gy + 0.5) * src->h / dh - 0.5;
            double bx = floor(tx), by = floor(ty);
            double fx = tx - bx, fy = ty - by;
            double acc[4] = {0}, wsum = 0;
            for (int y = 1 - bound; y <= bound; y++) {
                for (int x = 1 - bound; x <= bound; x++) {
                    double d = sqrt((x - fx) * (x - fx) + (y - fy) * (y - fy));
                    if (d >= radius)
                        continue;
                    double w = fn(priv, d);
                    const float *c = texel(src, (int) bx + x, (int) by + y);
                    wsum += w;
                    for (int k = 0; k < src->comps; k++)
                        acc[k] += w * c[k];
                }
            }
            for (int k = 0; k < src->comps; k++)
                dst[((size_t) gy * dw + gx) * src->comps + k] = acc[k] / wsum;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Array front-end to the portable math definitions (pl_oracle_math.h), for the tests that pin
// them against libm and against the CUDA restatement.
//   fn: 0 pow(a,b)  1 exp(a)  2 log(a)  3 sin(a)  4 cos(a)  5 atan2(a,b)  (binary64 definitions)
//       6 sin(a)  7 cos(a)  8 atan2(a,b)  (fp32 hot-path definitions)  9 log2(a)
//       16 + id: transfer curve `id` with its exact fallback (pl_oracle_color.c)
// ------------------------------------------------------------------------------------------
void orc_math(int fn, const float *a, const float *b, float *out, size_t n)
{
    if (fn >= 16) {
        orc_curve_apply(fn - 16, a, out, n);
        return;
    }
    #pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float s, c;
        switch (fn) {
        case 0: out[i] = pm_powf(a[i], b[i]); break;
        case 1: out[i] = pm_expf(a[i]); break;
        case 2: out[i] = pm_logf(a[i]); break;
        case 3: pm_sincosf(a[i], &s, &c); out[i] = s; break;
        case 4: pm_sincosf(a[i], &s, &c); out[i] = c; break;
        case 5: out[i] = pm_atan2f(a[i], b[i]); break;
        case 6: pf_sincosf(a[i], &s, &c); out[i] = s; break;
        case 7: pf_sincosf(a[i], &s, &c); out[i] = c; break;
        case 8: out[i] = pf_atan2f(a[i], b[i]); break;
        case 9: out[i] = pm_log2f(a[i]); break;
        default: out[i] = NAN;
        }
    }
}
