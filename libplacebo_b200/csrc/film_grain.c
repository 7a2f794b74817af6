// pl_shader_film_grain: AV1 grain synthesis (src/shaders/film_grain.c, film_grain_av1.c).
//
// Host side = the normative AV1 process run once per frame (film_grain_av1.c:303-518): a
// 73x82 Gaussian-noise template shaped by the auto-regressive filter (luma, then each chroma
// channel with its luma contribution), cropped to the 64x64 (or sub-sampled) LUT the GPU
// indexes, the per-32x32-block random offsets (with the left / top / top-left neighbours packed
// into the same word for the overlap blend) and the piecewise-linear scaling function as a
// 256-entry LUT. Everything is integer arithmetic: the tables are bit-identical to the
// reference's (tests/test_gpu_grain.py compares them with LUTs harvested from the real
// pl_shader_film_grain). The per-pixel part lives in pixel_core.cuh (plb_grain_av1).
#include "shader_priv.h"
#include "t0/av1_gaussian.h"
#include "t0/h274_tables.h"

enum {
    BLOCK_SIZE = 32,
    SCALING_LUT_SIZE = 256,
    GRAIN_WIDTH = 82,
    GRAIN_HEIGHT = 73,
    GRAIN_WIDTH_LUT = 64,
    GRAIN_HEIGHT_LUT = 64,
    GRAIN_PAD_LUT = 9,
    SUB_GRAIN_WIDTH = 44,
    SUB_GRAIN_HEIGHT = 38,
    SUB_GRAIN_PAD_LUT = 6,
};

enum { H274_RING = 8, H274_DB = 13 * 64, H274_MAX_IV = 3 * 256 };
#define H274_REC_SIZE (sizeof(struct plb_h274_plane) + H274_MAX_IV * sizeof(struct plb_h274_interval))

struct plb_grain_obj {
    // parameters the cached tables were generated for
    bool valid;
    struct pl_film_grain_data data;
    struct pl_color_repr repr;
    bool has[3];
    int sub_x, sub_y, off_w, off_h;
    // host scratch
    int16_t tmp_y[GRAIN_HEIGHT][GRAIN_WIDTH];
    int16_t tmp_uv[GRAIN_HEIGHT][GRAIN_WIDTH];
    float grain[3][GRAIN_HEIGHT_LUT * GRAIN_WIDTH_LUT];
    float scaling[3][SCALING_LUT_SIZE];
    uint32_t *offsets;
    // device copies
    float *dev_grain[3];
    float *dev_scaling[3];
    uint32_t *dev_offsets;
    size_t dev_offsets_size;
    // H.274: the grain data base and a small ring of per-pass parameter records
    float *dev_h274_db;
    struct plb_h274_plane *dev_h274[H274_RING];
    int h274_next;
};

void plb_grain_obj_free(pl_gpu gpu, struct plb_grain_obj *g)
{
    if (!g)
        return;
    for (int i = 0; i < 3; i++) {
        plb_dev_free(gpu, g->dev_grain[i]);
        plb_dev_free(gpu, g->dev_scaling[i]);
    }
    plb_dev_free(gpu, g->dev_offsets);
    plb_dev_free(gpu, g->dev_h274_db);
    for (int i = 0; i < H274_RING; i++)
        plb_dev_free(gpu, g->dev_h274[i]);
    free(g->offsets);
    free(g);
}

// film_grain.h:30-70: which colour channel does component i of the texture carry?
static enum pl_channel channel_map(int i, const struct pl_film_grain_params *params)
{
    if (i >= params->components)
        return PL_CHANNEL_NONE;
    const int comp = params->component_mapping[i];
    if (comp < 0 || comp > 2)
        return PL_CHANNEL_NONE;
    switch (params->repr->sys) {
    case PL_COLOR_SYSTEM_UNKNOWN:
    case PL_COLOR_SYSTEM_RGB: {
        static const enum pl_channel rgb[3] = { PL_CHANNEL_CR, PL_CHANNEL_Y, PL_CHANNEL_CB };
        return rgb[comp]; // R -> Cr, G -> Y, B -> Cb
    }
    case PL_COLOR_SYSTEM_XYZ: {
        static const enum pl_channel xyz[3] = { PL_CHANNEL_CR, PL_CHANNEL_Y, PL_CHANNEL_CB };
        return xyz[comp]; // X -> Cr, Y -> Y, Z -> Cb
    }
    default:
        return comp;
    }
}

static bool needs_av1(const struct pl_film_grain_params *params)
{
    const struct pl_av1_grain_data *data = &params->data.params.av1;
    const bool has_y = data->num_points_y > 0;
    const bool has_u = data->num_points_uv[0] > 0 || data->chroma_scaling_from_luma;
    const bool has_v = data->num_points_uv[1] > 0 || data->chroma_scaling_from_luma;
    for (int i = 0; i < 3; i++) {
        const enum pl_channel ch = channel_map(i, params);
        if ((ch == PL_CHANNEL_Y && has_y) || (ch == PL_CHANNEL_CB && has_u) ||
            (ch == PL_CHANNEL_CR && has_v))
            return true;
    }
    return false;
}

// film_grain_h274.c:127-143
static bool needs_h274(const struct pl_film_grain_params *params)
{
    const struct pl_h274_grain_data *data = &params->data.params.h274;
    if (data->model_id != 0)
        return false;
    for (int i = 0; i < 3; i++) {
        const enum pl_channel ch = channel_map(i, params);
        if (ch >= 0 && ch < 3 && data->component_model_present[ch])
            return true;
    }
    return false;
}

bool pl_needs_film_grain(const struct pl_film_grain_params *params)
{
    switch (params->data.type) {
    case PL_FILM_GRAIN_AV1:  return needs_av1(params);
    case PL_FILM_GRAIN_H274: return needs_h274(params);
    default:                 return false;
    }
}

// ---- the AV1 random process ---------------------------------------------------------------
static inline int get_random_number(int bits, uint16_t *state)
{
    const int r = *state;
    const uint16_t bit = ((r >> 0) ^ (r >> 1) ^ (r >> 3) ^ (r >> 12)) & 1;
    *state = (r >> 1) | (bit << 15);
    return (*state >> (16 - bits)) & ((1 << bits) - 1);
}

static inline int round2(int x, int shift)
{
    return shift ? (x + (1 << (shift - 1))) >> shift : x;
}

static inline int grain_bit_depth(const struct pl_color_repr *repr)
{
    const int depth = PLB_DEF(repr->bits.color_depth, PLB_DEF(repr->bits.sample_depth, 8));
    return PLB_MIN(depth, 12);
}

struct grain_scale {
    int grain_center, grain_min, grain_max;
    float texture_scale, grain_scale;
};

static struct grain_scale get_grain_scale(const struct pl_film_grain_params *params)
{
    const int bits = grain_bit_depth(params->repr);
    struct grain_scale ret = { .grain_center = 128 << (bits - 8) };
    ret.grain_min = -ret.grain_center;
    ret.grain_max = (256 << (bits - 8)) - 1 - ret.grain_center;
    struct pl_color_repr repr = *params->repr;
    ret.texture_scale = pl_color_repr_normalize(&repr);
    ret.grain_scale = 1.0 / ((1 << bits) - 1);
    return ret;
}

// LumaGrain of the specification
static void generate_grain_y(float *out, int16_t buf[GRAIN_HEIGHT][GRAIN_WIDTH],
                             const struct pl_film_grain_params *params)
{
    const struct pl_av1_grain_data *data = &params->data.params.av1;
    const struct grain_scale scale = get_grain_scale(params);
    uint16_t seed = (uint16_t) params->data.seed;
    const int shift = 12 - grain_bit_depth(params->repr) + data->grain_scale_shift;
    for (int y = 0; y < GRAIN_HEIGHT; y++) {
        for (int x = 0; x < GRAIN_WIDTH; x++)
            buf[y][x] = round2(plb_av1_gaussian_sequence[get_random_number(11, &seed)], shift);
    }
    const int pad = 3, lag = data->ar_coeff_lag;
    for (int y = pad; y < GRAIN_HEIGHT; y++) {
        for (int x = pad; x < GRAIN_WIDTH - pad; x++) {
            const int8_t *coeff = data->ar_coeffs_y;
            int sum = 0;
            for (int dy = -lag; dy <= 0; dy++) {
                for (int dx = -lag; dx <= lag; dx++) {
                    if (!dx && !dy)
                        break;
                    sum += *(coeff++) * buf[y + dy][x + dx];
                }
            }
            int16_t grain = buf[y][x] + round2(sum, data->ar_coeff_shift);
            buf[y][x] = PLB_CLAMP(grain, scale.grain_min, scale.grain_max);
        }
    }
    for (int y = 0; y < GRAIN_HEIGHT_LUT; y++) {
        for (int x = 0; x < GRAIN_WIDTH_LUT; x++)
            out[y * GRAIN_WIDTH_LUT + x] = buf[y + GRAIN_PAD_LUT][x + GRAIN_PAD_LUT] * scale.grain_scale;
    }
}

// CbGrain / CrGrain
static void generate_grain_uv(float *out, int16_t buf[GRAIN_HEIGHT][GRAIN_WIDTH],
                              const int16_t buf_y[GRAIN_HEIGHT][GRAIN_WIDTH],
                              enum pl_channel channel, int sub_x, int sub_y,
                              const struct pl_film_grain_params *params)
{
    const struct pl_av1_grain_data *data = &params->data.params.av1;
    const struct grain_scale scale = get_grain_scale(params);
    const int shift = 12 - grain_bit_depth(params->repr) + data->grain_scale_shift;
    uint16_t seed = params->data.seed;
    seed ^= channel == PL_CHANNEL_CB ? 0xb524 : 0x49d8;
    const int cw = sub_x ? SUB_GRAIN_WIDTH : GRAIN_WIDTH;
    const int ch = sub_y ? SUB_GRAIN_HEIGHT : GRAIN_HEIGHT;
    const int8_t *coeffs = data->ar_coeffs_uv[channel == PL_CHANNEL_CB ? 0 : 1];
    for (int y = 0; y < ch; y++) {
        for (int x = 0; x < cw; x++)
            buf[y][x] = round2(plb_av1_gaussian_sequence[get_random_number(11, &seed)], shift);
    }
    const int pad = 3, lag = data->ar_coeff_lag;
    for (int y = pad; y < ch; y++) {
        for (int x = pad; x < cw - pad; x++) {
            const int8_t *coeff = coeffs;
            int sum = 0;
            for (int dy = -lag; dy <= 0; dy++) {
                for (int dx = -lag; dx <= lag; dx++) {
                    if (!dx && !dy) {
                        // the current position takes the co-sited luma grain instead
                        if (!data->num_points_y)
                            break;
                        int luma = 0;
                        const int lx = ((x - pad) << sub_x) + pad, ly = ((y - pad) << sub_y) + pad;
                        for (int i = 0; i <= sub_y; i++) {
                            for (int j = 0; j <= sub_x; j++)
                                luma += buf_y[ly + i][lx + j];
                        }
                        luma = round2(luma, sub_x + sub_y);
                        sum += luma * (*coeff);
                        break;
                    }
                    sum += *(coeff++) * buf[y + dy][x + dx];
                }
            }
            int16_t grain = buf[y][x] + round2(sum, data->ar_coeff_shift);
            buf[y][x] = PLB_CLAMP(grain, scale.grain_min, scale.grain_max);
        }
    }
    const int lw = GRAIN_WIDTH_LUT >> sub_x, lh = GRAIN_HEIGHT_LUT >> sub_y;
    const int px = sub_x ? SUB_GRAIN_PAD_LUT : GRAIN_PAD_LUT, py = sub_y ? SUB_GRAIN_PAD_LUT : GRAIN_PAD_LUT;
    for (int y = 0; y < lh; y++) {
        for (int x = 0; x < lw; x++)
            out[y * lw + x] = buf[y + py][x + px] * scale.grain_scale;
    }
}

static void generate_offsets(uint32_t *buf, int w, int h, uint64_t seed64)
{
    for (int y = 0; y < h; y++) {
        uint16_t state = seed64;
        state ^= ((y * 37 + 178) & 0xFF) << 8;
        state ^= ((y * 173 + 105) & 0xFF);
        for (int x = 0; x < w; x++) {
            uint32_t *o = &buf[y * w + x];
            const uint8_t val = get_random_number(8, &state);
            const uint8_t val_l = x ? (uint8_t) o[-1] : 0;
            const uint8_t val_t = y ? (uint8_t) o[-w] : 0;
            const uint8_t val_tl = x && y ? (uint8_t) o[-w - 1] : 0;
            *o = ((uint32_t) val_tl << 24) | ((uint32_t) val_t << 16) | ((uint32_t) val_l << 8) | val;
        }
    }
}

static void generate_scaling(float *out, int num, const uint8_t (*points)[2], int scaling_shift)
{
    const float range = 1 << scaling_shift;
    for (int i = 0; i < points[0][0]; i++)
        out[i] = points[0][1] / range;
    for (int i = 0; i < num - 1; i++) {
        const int bx = points[i][0], by = points[i][1];
        const int dx = points[i + 1][0] - bx, dy = points[i + 1][1] - by;
        const int delta = dy * ((0x10000 + (dx >> 1)) / dx);
        for (int x = 0; x < dx; x++)
            out[bx + x] = (by + ((x * delta + 0x8000) >> 16)) / range;
    }
    for (int i = points[num - 1][0]; i < SCALING_LUT_SIZE; i++)
        out[i] = points[num - 1][1] / range;
}

static bool av1_data_eq(const struct pl_film_grain_data *da, const struct pl_film_grain_data *db)
{
    const struct pl_av1_grain_data *a = &da->params.av1, *b = &db->params.av1;
    return da->seed == db->seed && a->chroma_scaling_from_luma == b->chroma_scaling_from_luma &&
           a->scaling_shift == b->scaling_shift && a->ar_coeff_lag == b->ar_coeff_lag &&
           a->ar_coeff_shift == b->ar_coeff_shift && a->grain_scale_shift == b->grain_scale_shift &&
           a->num_points_y == b->num_points_y && !memcmp(a->points_y, b->points_y, sizeof(a->points_y)) &&
           !memcmp(a->num_points_uv, b->num_points_uv, sizeof(a->num_points_uv)) &&
           !memcmp(a->points_uv, b->points_uv, sizeof(a->points_uv)) &&
           !memcmp(a->ar_coeffs_y, b->ar_coeffs_y, sizeof(a->ar_coeffs_y)) &&
           !memcmp(a->ar_coeffs_uv, b->ar_coeffs_uv, sizeof(a->ar_coeffs_uv));
}

static bool upload_f(pl_gpu gpu, float **dev, const float *host, size_t n)
{
    if (!*dev)
        *dev = plb_dev_alloc(gpu, n * sizeof(float));
    return *dev && plb_dev_upload(gpu, *dev, host, n * sizeof(float));
}

static bool shader_fg_av1(pl_shader sh, struct plb_grain_obj *obj,
                          const struct pl_film_grain_params *params)
{
    pl_gpu gpu = sh->params.gpu;
    const int tex_w = params->tex->params.w, tex_h = params->tex->params.h;
    int sub_x = 0, sub_y = 0;
    if (params->luma_tex) {
        sub_x = params->luma_tex->params.w > tex_w;
        sub_y = params->luma_tex->params.h > tex_h;
    }
    const struct pl_av1_grain_data *data = &params->data.params.av1;
    if (grain_bit_depth(params->repr) < 8 || 12 - grain_bit_depth(params->repr) + data->grain_scale_shift < 0 ||
        data->ar_coeff_lag < 0 || data->ar_coeff_lag > 3) {
        plb_shader_fail(sh, "AV1 film grain: invalid bit depth / parameters");
        return false;
    }
    bool has[3] = { data->num_points_y > 0,
                    data->num_points_uv[0] > 0 || data->chroma_scaling_from_luma,
                    data->num_points_uv[1] > 0 || data->chroma_scaling_from_luma };
    bool tex_is[3] = { false, false, false };
    for (int i = 0; i < 3; i++) {
        const enum pl_channel ch = channel_map(i, params);
        if (ch >= 0 && ch < 3)
            tex_is[ch] = true;
    }
    if (tex_is[0] && (sub_x || sub_y)) {
        PLB_WARN(sh->log, "pl_film_grain_params.channels includes PL_CHANNEL_Y but plane is "
                 "subsampled, this makes no sense. Continuing anyway but output is likely "
                 "incorrect.");
    }
    if (sh->failed || sh->src != PLB_SRC_NONE) {
        plb_shader_fail(sh, "pl_shader_film_grain: shader must be empty (sh_require PL_SHADER_SIG_NONE)");
        return false;
    }
    for (int c = 0; c < 3; c++)
        has[c] &= tex_is[c];

    const int off_w = ((((tex_w << sub_x) + 127) & ~127)) / 32, off_h = ((((tex_h << sub_y) + 127) & ~127)) / 32;
    const bool update = !obj->valid || !av1_data_eq(&params->data, &obj->data) ||
                        !pl_color_repr_equal(params->repr, &obj->repr) ||
                        memcmp(has, obj->has, sizeof(has)) || obj->sub_x != sub_x ||
                        obj->sub_y != sub_y || obj->off_w != off_w || obj->off_h != off_h;
    if (update) {
        bool ok = true;
        generate_grain_y(obj->grain[0], obj->tmp_y, params); // needed even for chroma only
        if (has[0])
            ok &= upload_f(gpu, &obj->dev_grain[0], obj->grain[0], GRAIN_WIDTH_LUT * GRAIN_HEIGHT_LUT);
        for (int c = 1; c < 3; c++) {
            if (!has[c])
                continue;
            generate_grain_uv(obj->grain[c], obj->tmp_uv, obj->tmp_y, c, sub_x, sub_y, params);
            ok &= upload_f(gpu, &obj->dev_grain[c], obj->grain[c],
                           (GRAIN_WIDTH_LUT >> sub_x) * (GRAIN_HEIGHT_LUT >> sub_y));
        }
        const size_t osz = (size_t) off_w * off_h * sizeof(uint32_t);
        if (osz > obj->dev_offsets_size) {
            plb_dev_free(gpu, obj->dev_offsets);
            free(obj->offsets);
            obj->dev_offsets = plb_dev_alloc(gpu, osz);
            obj->offsets = malloc(osz);
            obj->dev_offsets_size = osz;
        }
        ok &= obj->offsets && obj->dev_offsets;
        if (ok) {
            generate_offsets(obj->offsets, off_w, off_h, params->data.seed);
            ok &= plb_dev_upload(gpu, obj->dev_offsets, obj->offsets, osz);
        }
        for (int c = 0; c < 3 && ok; c++) {
            int num;
            const uint8_t (*points)[2];
            if (c == 0 || data->chroma_scaling_from_luma) {
                num = data->num_points_y;
                points = &data->points_y[0];
            } else {
                num = data->num_points_uv[c - 1];
                points = &data->points_uv[c - 1][0];
            }
            if (has[c] && num > 0) {
                generate_scaling(obj->scaling[c], num, points, data->scaling_shift);
                ok &= upload_f(gpu, &obj->dev_scaling[c], obj->scaling[c], SCALING_LUT_SIZE);
            }
        }
        if (!ok) {
            plb_shader_fail(sh, "Failed generating/uploading film grain LUTs!");
            obj->valid = false;
            return false;
        }
        obj->valid = true;
        obj->data = params->data;
        obj->repr = *params->repr;
        memcpy(obj->has, has, sizeof(has));
        obj->sub_x = sub_x;
        obj->sub_y = sub_y;
        obj->off_w = off_w;
        obj->off_h = off_h;
    }

    // ---- record the plane pass
    const struct grain_scale scale = get_grain_scale(params);
    pl_color_repr_normalize(params->repr); // the reference updates the caller's repr (:866)
    const int bits = PLB_DEF(params->repr->bits.color_depth, 8);
    sh->src = PLB_SRC_PLANES;
    sh->is_compute = true;
    sh->wg_w = BLOCK_SIZE >> sub_x;
    sh->wg_h = BLOCK_SIZE >> sub_y;
    sh->out_w = tex_w;
    sh->out_h = tex_h;
    plb_shader_describe(sh, "AV1 film grain");
    plb_shader_use_tex(sh, params->tex);
    struct plb_plane_src *pl = &sh->px.planes[0];
    memset(pl, 0, sizeof(*pl));
    sh->px.nplanes = 1;
    sh->px.init[3] = 1.0f;
    pl->kind = PLB_PLANE_GRAIN_AV1;
    pl->tex = plb_tex_priv(params->tex)->d;
    pl->rect = (pl_rect2df) { 0, 0, tex_w, tex_h };
    pl->scale = 1.0f;
    for (int k = 0; k < 4; k++)
        pl->map[k] = k;
    struct plb_av1_plane *g = &pl->av1;
    g->ncomp = params->components;
    g->offsets = obj->dev_offsets;
    g->off_w = off_w;
    g->sub_x = sub_x;
    g->sub_y = sub_y;
    g->overlap = data->overlap;
    g->tex_scale = scale.texture_scale;
    g->grain_min = scale.grain_min * scale.grain_scale;
    g->grain_max = scale.grain_max * scale.grain_scale;
    if (pl_color_levels_guess(params->repr) == PL_COLOR_LEVELS_LIMITED) {
        const float out_scale = (1 << bits) / ((1 << bits) - 1.0);
        g->min_value = 16 / 256.0 * out_scale;
        g->max_luma = 235 / 256.0 * out_scale;
        g->max_chroma = 240 / 256.0 * out_scale;
        if (!pl_color_system_is_ycbcr_like(params->repr->sys))
            g->max_chroma = g->max_luma;
    } else {
        g->min_value = 0.0f;
        g->max_luma = g->max_chroma = 1.0f;
    }
    g->chroma_from_luma = data->chroma_scaling_from_luma;
    g->luma_from_comp = -2;
    for (int i = 0; i < 3; i++) {
        const enum pl_channel ch = channel_map(i, params);
        g->channel[i] = ch;
        if (ch < 0 || ch > 2)
            continue;
        int num = (ch == 0 || data->chroma_scaling_from_luma) ? data->num_points_y
                                                               : data->num_points_uv[ch - 1];
        if (has[ch] && num > 0) {
            g->grain[i] = obj->dev_grain[ch];
            g->grain_w[i] = ch == 0 ? GRAIN_WIDTH_LUT : GRAIN_WIDTH_LUT >> sub_x;
            g->scaling[i] = obj->dev_scaling[ch];
        }
        if (ch != PL_CHANNEL_Y) {
            g->mult[i][0] = data->uv_mult_luma[ch - 1] / 64.0;
            g->mult[i][1] = data->uv_mult[ch - 1] / 64.0;
            const int c_offset = (unsigned) data->uv_offset[ch - 1] << (bits - 8);
            g->offset[i] = c_offset * scale.grain_scale;
        }
    }
    if (tex_is[1] || tex_is[2]) {
        if (tex_is[0]) {
            for (int i = 0; i < 3; i++) {
                if (channel_map(i, params) == PL_CHANNEL_Y) {
                    g->luma_from_comp = i;
                    break;
                }
            }
        } else {
            if (!params->luma_tex) {
                plb_shader_fail(sh, "AV1 film grain: chroma planes need pl_film_grain_params.luma_tex");
                return false;
            }
            g->luma_from_comp = -1;
            g->luma_comp = params->luma_comp;
            pl->luma = plb_tex_priv(params->luma_tex)->d;
            plb_shader_use_tex(sh, params->luma_tex);
        }
    }
    sh->px.coord_mode = PLB_COORD_EXACT;
    return true;
}


// ---- H.274 / SMPTE RDD 5 (film_grain_h274.c) ------------------------------------------------
// The grain data base: for every pair of cut-off frequencies (h, v) in 0..12 a 64x64 block of
// band-limited Gaussian noise: LUT noise in the (4h+12)x(4v+12) low-frequency corner, inverse
// 64x64 integer transform (R64T), horizontal-edge attenuation. Integer arithmetic throughout;
// the final value is int / 255.0 rounded to float (film_grain_h274.c:33-105).
static void h274_slice(float *out, size_t stride, int h, int v, int8_t (*grain)[64],
                       int16_t (*tmp)[64])
{
    const int freq_h = ((h + 3) << 2) - 1, freq_v = ((v + 3) << 2) - 1;
    uint32_t lfsr = plb_h274_seed[h + v * 13];
    // column-major start image: grain[x][y]
    for (int y = 0; y <= freq_v; y++) {
        for (int x = 0; x <= freq_h; x += 4) {
            const int at = lfsr % 2048;
            for (int k = 0; k < 4; k++)
                grain[x + k][y] = plb_h274_gaussian[at + k];
            // x^31 + x^3 + 1
            const uint32_t fb = 1u ^ (lfsr >> 2) ^ (lfsr >> 30);
            lfsr = (lfsr << 1) | (fb & 1u);
        }
    }
    grain[0][0] = 0;
    for (int y = 0; y < 64; y++) {
        for (int x = 0; x <= freq_h; x++) {
            int32_t acc = 0;
            for (int p = 0; p <= freq_v; p++)
                acc += plb_h274_r64t[y][p] * grain[x][p];
            tmp[y][x] = (acc + 128) >> 8;
        }
    }
    for (int y = 0; y < 64; y++) {
        for (int x = 0; x < 64; x++) {
            int32_t acc = 0;
            for (int p = 0; p <= freq_h; p++)
                acc += tmp[y][p] * plb_h274_r64t[x][p];
            acc = (acc + 128) >> 8;
            grain[y][x] = PLB_CLAMP(acc, -127, 127);
        }
    }
    static const uint8_t attenuate[13] = { 64, 71, 77, 84, 90, 96, 103, 109, 116, 122, 128, 128, 128 };
    for (int y = 0; y < 64; y++, out += stride) {
        const bool edge = (y & 7) == 0 || (y & 7) == 7;
        for (int x = 0; x < 64; x++) {
            const int g = edge ? (grain[y][x] * attenuate[v]) >> 7 : grain[y][x];
            out[x] = g / 255.0;
        }
    }
}

// fill_grain_lut (film_grain_h274.c:107-125): slice (h, v) starts at row 64h, column 64v
static void h274_database(float *db)
{
    int8_t (*grain)[64] = malloc(64 * 64);
    int16_t (*tmp)[64] = malloc(64 * 64 * sizeof(int16_t));
    for (int h = 0; h < 13 && grain && tmp; h++) {
        for (int v = 0; v < 13; v++)
            h274_slice(db + (size_t) (h * 64) * H274_DB + v * 64, H274_DB, h, v, grain, tmp);
    }
    free(grain);
    free(tmp);
}

bool pl_cuda_h274_database_host(float *out)
{
    if (!out)
        return false;
    h274_database(out);
    return true;
}

// pl_shader_fg_h274 (film_grain_h274.c:145-315)
static bool shader_fg_h274(pl_shader sh, struct plb_grain_obj *obj,
                           const struct pl_film_grain_params *params)
{
    pl_gpu gpu = sh->params.gpu;
    const struct pl_h274_grain_data *data = &params->data.params.h274;
    const int tex_w = params->tex->params.w, tex_h = params->tex->params.h;
    if (!obj->dev_h274_db) {
        float *db = malloc(sizeof(float) * H274_DB * H274_DB);
        if (!db)
            return false;
        h274_database(db);
        const bool ok = upload_f(gpu, &obj->dev_h274_db, db, (size_t) H274_DB * H274_DB);
        free(db);
        if (!ok) {
            plb_shader_fail(sh, "H.274 film grain: uploading the grain data base failed");
            return false;
        }
    }

    struct plb_h274_plane *rec = calloc(1, H274_REC_SIZE);
    if (!rec)
        return false;
    rec->db = obj->dev_h274_db;
    rec->seed = (unsigned int) params->data.seed;
    rec->tex_scale = pl_color_repr_normalize(params->repr);
    rec->scale_factor = 1.0 / (1 << (data->log2_scale_factor + 6));
    int niv = 0;
    for (int idx = 0; idx < params->components; idx++) {
        const enum pl_channel c = channel_map(idx, params);
        if (c == PL_CHANNEL_NONE || !data->component_model_present[c])
            continue;
        const int e = rec->nent++;
        rec->ent[e].channel = c;
        rec->ent[e].first = niv;
        const int count = PLB_MIN((int) data->num_intensity_intervals[c], 256);
        for (int i = 0; i < count; i++, niv++) {
            struct plb_h274_interval *iv = &rec->iv[niv];
            iv->lo = data->intensity_interval_lower_bound[c][i] / 255.0;
            iv->hi = data->intensity_interval_upper_bound[c][i] / 255.0;
            const uint8_t nv = data->num_model_values[c];
            uint8_t h = nv > 1 ? data->comp_model_value[c][i][1] : 8;
            uint8_t v = nv > 2 ? data->comp_model_value[c][i][2] : h;
            h = PLB_CLAMP(h, 2, 14) - 2;
            v = PLB_CLAMP(v, 2, 14) - 2;
            int16_t scale = data->comp_model_value[c][i][0];
            if (c > 0 && pl_color_system_is_ycbcr_like(params->repr->sys))
                scale >>= 1; // chroma gets half the strength
            iv->value = (uint32_t) (uint16_t) scale << 16 | h << 8 | v;
        }
        rec->ent[e].count = count;
    }
    const int slot = obj->h274_next++ % H274_RING;
    if (!obj->dev_h274[slot])
        obj->dev_h274[slot] = plb_dev_alloc(gpu, H274_REC_SIZE);
    const size_t used = sizeof(*rec) + niv * sizeof(struct plb_h274_interval);
    const bool ok = obj->dev_h274[slot] && plb_dev_upload(gpu, obj->dev_h274[slot], rec, used);
    free(rec);
    if (!ok) {
        plb_shader_fail(sh, "H.274 film grain: uploading the parameters failed");
        return false;
    }

    sh->src = PLB_SRC_PLANES;
    sh->is_compute = true;
    sh->wg_w = sh->wg_h = 8;
    sh->out_w = tex_w;
    sh->out_h = tex_h;
    plb_shader_describe(sh, "H.274 film grain");
    plb_shader_use_tex(sh, params->tex);
    struct plb_plane_src *pl = &sh->px.planes[0];
    memset(pl, 0, sizeof(*pl));
    sh->px.nplanes = 1;
    sh->px.init[3] = 1.0f;
    pl->kind = PLB_PLANE_GRAIN_H274;
    pl->tex = plb_tex_priv(params->tex)->d;
    pl->rect = (pl_rect2df) { 0, 0, tex_w, tex_h };
    pl->scale = 1.0f;
    for (int k = 0; k < 4; k++)
        pl->map[k] = k;
    pl->h274 = obj->dev_h274[slot];
    sh->px.coord_mode = PLB_COORD_EXACT;
    return true;
}

bool pl_shader_film_grain(pl_shader sh, pl_shader_obj *grain_state,
                          const struct pl_film_grain_params *params)
{
    if (!params || !params->tex || !params->repr) {
        plb_shader_fail(sh, "pl_shader_film_grain: tex and repr are required");
        return false;
    }
    if (!pl_needs_film_grain(params)) {
        plb_shader_fail(sh, "pl_shader_film_grain called but no film grain needs to be applied, "
                        "test with `pl_needs_film_grain` first!");
        return false;
    }
    struct plb_grain_obj **pobj = (struct plb_grain_obj **) plb_shader_obj(sh, grain_state, PLB_OBJ_GRAIN);
    if (!pobj)
        return false;
    if (!*pobj)
        *pobj = calloc(1, sizeof(**pobj));
    if (!*pobj)
        return false;
    if (params->data.type == PL_FILM_GRAIN_H274)
        return shader_fg_h274(sh, *pobj, params);
    return shader_fg_av1(sh, *pobj, params);
}

// renderer.c:1355-1429 `plane_film_grain`
bool plb_plane_film_grain(pl_renderer rr, pl_shader sh, pl_shader_obj *state,
                          const struct pl_frame *image, const struct pl_plane *plane,
                          pl_tex tex, pl_tex luma_tex, int luma_comp, struct pl_color_repr *repr)
{
    (void) rr;
    struct pl_film_grain_params gp = {
        .data = image->film_grain,
        .tex = tex,
        .repr = repr,
        .components = plane->components,
        .luma_tex = luma_tex,
        .luma_comp = luma_comp,
    };
    for (int c = 0; c < plane->components; c++)
        gp.component_mapping[c] = plane->component_mapping[c];
    if (!pl_needs_film_grain(&gp))
        return false;
    return pl_shader_film_grain(sh, state, &gp);
}

// Diagnostics (declared in the public header): the host-generated AV1 tables for the given
// parameters, without touching a GPU. Arrays: grain_y[64*64], grain_cb / grain_cr
// [(64 >> sub_x) * (64 >> sub_y)], offsets[off_w * off_h] with off = align(dim << sub, 128) / 32,
// scaling[3][256]. Returns false for invalid parameters.
bool pl_cuda_av1_tables_host(const struct pl_film_grain_data *data, const struct pl_color_repr *repr,
                             int sub_x, int sub_y, int tex_w, int tex_h, float *grain_y,
                             float *grain_cb, float *grain_cr, uint32_t *offsets, float *scaling)
{
    if (!data || data->type != PL_FILM_GRAIN_AV1 || !repr)
        return false;
    struct pl_color_repr r = *repr;
    struct pl_film_grain_params params = { .data = *data, .repr = &r, .components = 3,
                                           .component_mapping = {0, 1, 2} };
    const struct pl_av1_grain_data *av1 = &data->params.av1;
    if (grain_bit_depth(&r) < 8 || 12 - grain_bit_depth(&r) + av1->grain_scale_shift < 0)
        return false;
    static int16_t tmp_y[GRAIN_HEIGHT][GRAIN_WIDTH], tmp_uv[GRAIN_HEIGHT][GRAIN_WIDTH];
    float gy[GRAIN_WIDTH_LUT * GRAIN_HEIGHT_LUT];
    generate_grain_y(gy, tmp_y, &params);
    if (grain_y)
        memcpy(grain_y, gy, sizeof(gy));
    if (grain_cb)
        generate_grain_uv(grain_cb, tmp_uv, tmp_y, PL_CHANNEL_CB, sub_x, sub_y, &params);
    if (grain_cr)
        generate_grain_uv(grain_cr, tmp_uv, tmp_y, PL_CHANNEL_CR, sub_x, sub_y, &params);
    if (offsets) {
        const int off_w = (((tex_w << sub_x) + 127) & ~127) / 32, off_h = (((tex_h << sub_y) + 127) & ~127) / 32;
        generate_offsets(offsets, off_w, off_h, data->seed);
    }
    if (scaling) {
        for (int c = 0; c < 3; c++) {
            int num = (c == 0 || av1->chroma_scaling_from_luma) ? av1->num_points_y : av1->num_points_uv[c - 1];
            const uint8_t (*points)[2] = (c == 0 || av1->chroma_scaling_from_luma)
                ? &av1->points_y[0] : &av1->points_uv[c - 1][0];
            if (num > 0)
                generate_scaling(scaling + 256 * c, num, points, av1->scaling_shift);
        }
    }
    return true;
}
