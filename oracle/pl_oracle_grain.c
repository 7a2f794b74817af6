// TEST INFRASTRUCTURE — CPU oracle, never linked into the product.
//
// pl_shader_fg_h274 (src/shaders/film_grain_h274.c:145-315) is restated at the end of the file.
// pl_shader_fg_av1, the per-pixel GLSL of src/shaders/film_grain_av1.c:827-1013, restated in C.
// The LUTs (grain templates, block offsets, scaling functions) are INPUTS: tests feed the ones
// harvested from the real reference shader (oracle/harvest.c hvs_av1_grain), so that only the
// pixel arithmetic is under test here.
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#include "pl_oracle.h"

static inline float clampf_(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// sh_lut linear lookup of a 1-D float LUT (lut.c:285-295 + 700-715 in array form)
static inline float lut1d(const float *lut, int n, float x)
{
    float fpos = clampf_(x, 0.0f, 1.0f) * (float) (n - 1);
    float fbase = floorf(fpos), fceil = ceilf(fpos);
    float f = fpos - fbase;
    return lut[(int) fbase] * (1.0f - f) + lut[(int) fceil] * f;
}

// `sample(sh, off, lut, idx, sub_x, sub_y)` :499-518
static inline float sample_grain(const struct orc_av1_grain_params *p, int comp, uint32_t data,
                                 int off, int lx, int ly)
{
    const int dx = (off & 8) ? 1 : 0, dy = (off & 16) ? 1 : 0;
    const int bw = 32 >> p->sub_x, bh = 32 >> p->sub_y;
    const int ox = (p->sub_x ? 1 : 2) * (int) ((data >> (off + 4)) & 0xFu);
    const int oy = (p->sub_y ? 1 : 2) * (int) ((data >> off) & 0xFu);
    const int px = ox + lx + bw * dx, py = oy + ly + bh * dy;
    return p->grain[comp][py * p->grain_w + px];
}

void orc_av1_grain(const struct orc_tex *plane, const struct orc_tex *luma,
                   const struct orc_av1_grain_params *p, float *dst)
{
    const int w = plane->w, h = plane->h;
    const int bw = 32 >> p->sub_x, bh = 32 >> p->sub_y;
    const float bits_scale = 1.0f / (float) ((1 << p->bit_depth) - 1);
    const int center = 128 << (p->bit_depth - 8);
    const float grain_min = (float) (-center) * bits_scale;
    const float grain_max = (float) ((256 << (p->bit_depth - 8)) - 1 - center) * bits_scale;
    const float w_x[2] = { p->sub_x ? 23.0f / 32.0f : 27.0f / 32.0f, p->sub_x ? 22.0f / 32.0f : 17.0f / 32.0f };
    const float w_y[2] = { p->sub_y ? 23.0f / 32.0f : 27.0f / 32.0f, p->sub_y ? 22.0f / 32.0f : 17.0f / 32.0f };

    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < h; gy++) {
        for (int gx = 0; gx < w; gx++) {
            const int bx = gx / bw, by = gy / bh, lx = gx - bw * bx, ly = gy - bh * by;
            const uint32_t data = p->offsets[by * p->offsets_w + bx];
            const float *t = plane->data + ((size_t) gy * w + gx) * plane->comps;
            float color[4] = { 0, 0, 0, 1 }; // texelFetch fills missing components with (0,0,0,1)
            for (int k = 0; k < plane->comps; k++)
                color[k] = t[k];
            for (int k = 0; k < 4; k++)
                color[k] = p->scale * color[k];

            float average_luma = 0.0f;
            bool have_chroma = false;
            int y_comp = -1;
            for (int i = 0; i < p->num_comps; i++) {
                have_chroma |= !p->is_luma[i];
                if (p->is_luma[i] && y_comp < 0)
                    y_comp = i;
            }
            if (have_chroma) {
                if (y_comp >= 0) {
                    average_luma = color[y_comp];
                } else {
                    const int px = gx << p->sub_x, py = gy << p->sub_y;
                    average_luma = luma->data[((size_t) py * luma->w + px) * luma->comps + p->luma_comp];
                    average_luma *= p->scale;
                }
            }

            for (int i = 0; i < p->num_comps; i++) {
                if (!p->scaling[i])
                    continue;
                float grain = sample_grain(p, i, data, 0, lx, ly);
                if (p->overlap) {
                    const bool ovx = bx > 0 && lx < (2 >> p->sub_x);
                    if (ovx) {
                        // w = weights / 32; if (local_id.x == 1) w = w.yx; dot(vec2(val, grain), w)
                        const float wa = lx == 1 ? w_x[1] : w_x[0], wb = lx == 1 ? w_x[0] : w_x[1];
                        const float val = sample_grain(p, i, data, 8, lx, ly);
                        grain = fmaf(grain, wb, val * wa);
                    }
                    if (by > 0 && ly < (2 >> p->sub_y)) {
                        const float wa = ly == 1 ? w_y[1] : w_y[0], wb = ly == 1 ? w_y[0] : w_y[1];
                        float val;
                        if (ovx) {
                            const float va = lx == 1 ? w_x[1] : w_x[0], vb = lx == 1 ? w_x[0] : w_x[1];
                            const float tmp = sample_grain(p, i, data, 24, lx, ly);
                            val = sample_grain(p, i, data, 16, lx, ly);
                            val = fmaf(val, vb, tmp * va);
                        } else {
                            val = sample_grain(p, i, data, 16, lx, ly);
                        }
                        grain = fmaf(grain, wb, val * wa);
                    }
                    grain = clampf_(grain, grain_min, grain_max);
                }
                if (p->is_luma[i]) {
                    color[i] = fmaf(lut1d(p->scaling[i], 256, color[i]), grain, color[i]);
                    color[i] = clampf_(color[i], p->minValue, p->maxLuma);
                } else {
                    float val = average_luma;
                    if (!p->chroma_scaling_from_luma) {
                        val = fmaf(color[i], p->uv_mult[i], val * p->uv_mult_luma[i]);
                        val += p->uv_offset[i];
                    }
                    color[i] = fmaf(lut1d(p->scaling[i], 256, val), grain, color[i]);
                    color[i] = clampf_(color[i], p->minValue, p->maxChroma);
                }
            }
            float *o = dst + ((size_t) gy * w + gx) * 4;
            for (int k = 0; k < 4; k++)
                o[k] = color[k];
        }
    }
}


// ---- H.274 / SMPTE RDD 5 (src/shaders/film_grain_h274.c:145-315) ---------------------------
// The shader runs 8x8 work groups on a device with 32-wide subgroups (NVIDIA: subgroup k holds
// local invocation indices 32k .. 32k+31, i.e. rows 4k .. 4k+3 of the group). Per component:
//   avg = color[c] / 64; avg = subgroupAdd(avg); one invocation per subgroup does
//   atomicAdd(group_sum, int(avg * 10000000.0)); avg = float(group_sum) / 10000000.0   (:232-246)
// GLSL leaves the association order of subgroupAdd open; this restatement fixes it as ascending
// invocation index. Invocations outside the texture read (0,0,0,0) (robust access).
// `db` is the 832x832 grain data base (harvested from the reference's sh_lut fill, :107-125),
// fetched as db(offset) = texel (x = offset.x, y = offset.y).
void orc_h274_grain(const struct orc_tex *plane, const struct orc_h274_grain_params *p, float *dst)
{
    const int w = plane->w, h = plane->h;
    #pragma omp parallel for schedule(static)
    for (int gy = 0; gy < h; gy++) {
        for (int gx = 0; gx < w; gx++) {
            const int wx = gx / 8, wy = gy / 8;
            const float *t = plane->data + ((size_t) gy * w + gx) * plane->comps;
            float color[4] = { 0, 0, 0, 1 };
            for (int k = 0; k < plane->comps; k++)
                color[k] = t[k];
            for (int k = 0; k < 4; k++)
                color[k] = p->scale * color[k];
            // pcg3d :197-206
            uint32_t pcg[3] = { p->seed, (uint32_t) wx / 2u, (uint32_t) wy / 2u };
            for (int k = 0; k < 3; k++)
                pcg[k] = pcg[k] * 1664525u + 1013904223u;
            pcg[0] += pcg[1] * pcg[2];
            pcg[1] += pcg[2] * pcg[0];
            pcg[2] += pcg[0] * pcg[1];
            for (int k = 0; k < 3; k++)
                pcg[k] ^= pcg[k] >> 16;
            pcg[0] += pcg[1] * pcg[2];
            pcg[1] += pcg[2] * pcg[0];
            pcg[2] += pcg[0] * pcg[1];

            for (int e = 0; e < p->num_entries; e++) {
                const int c = p->channel[e];
                int group_sum = 0;
                for (int sg = 0; sg < 2; sg++) {
                    float sum = 0.0f;
                    for (int li = 32 * sg; li < 32 * sg + 32; li++) {
                        const int x = 8 * wx + (li & 7), y = 8 * wy + (li >> 3);
                        float v[4] = { 0, 0, 0, 0 };
                        if (x < w && y < h) {
                            const float *q = plane->data + ((size_t) y * w + x) * plane->comps;
                            v[3] = 1.0f;
                            for (int k = 0; k < plane->comps; k++)
                                v[k] = q[k];
                        }
                        sum += (p->scale * v[c]) / 64.0f;
                    }
                    group_sum += (int) (sum * 10000000.0f);
                }
                const float avg = (float) group_sum / 10000000.0f;
                uint32_t val = 0u;
                for (int i = 0; i < p->num_intervals[e]; i++) {
                    if (avg >= p->lower[e][i] && avg <= p->upper[e][i]) {
                        val = p->value[e][i];
                        break;
                    }
                }
                uint32_t ox = (val & 0xFF00u) >> 2, oy = (val & 0xFFu) << 6;
                float scale = p->scale_factor * (float) (int) (val >> 16);
                const uint32_t rnd = pcg[c];
                ox += (rnd >> 16) % 52u;
                oy += (rnd & 0xFFFFu) % 56u;
                ox &= 0xFFFCu;
                oy &= 0xFFF8u;
                if ((rnd & 1u) == 1u)
                    scale = -scale;
                ox += 8u * ((uint32_t) wx % 2u) + (uint32_t) (gx % 8);
                oy += 8u * ((uint32_t) wy % 2u) + (uint32_t) (gy % 8);
                const float grain = p->db[(size_t) oy * p->db_w + ox];
                color[c] = fmaf(scale, grain, color[c]);
            }
            for (int k = 0; k < 4; k++)
                dst[((size_t) gy * w + gx) * 4 + k] = color[k];
        }
    }
}
