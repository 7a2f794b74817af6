// Host-side colour-space parameter derivation: matrices, nominal luminance, inference.
// Behavioural restatement of the reference's src/colorspace.c and the matrix helpers of
// src/common.c, written from the formulas; every float operation keeps the reference's
// operand order and precision (float vs double) so that results are bit-identical
// (pinned by tests/test_t0_colorspace.py against oracle/_ref).
//   pl_color_repr_normalize      colorspace.c:190-216
//   pl_hdr_rescale               colorspace.c:367-418
//   pl_color_space_nominal_luma_ex :790-859, pl_color_space_infer(_map) :861-972
//   primaries table              :1117-1241, rgb2xyz :1243-1281, CAT16 adaptation :1289-1320
//   IPT matrices                 :1351-1390, primaries clip :1610-1667
//   pl_color_repr_decode         :1717-1896
#define _GNU_SOURCE
#include "../internal.h"

#include <math.h>

// ---- 3x3 helpers (src/common.c:100-245) -------------------------------------------------------

const pl_matrix3x3 pl_matrix3x3_identity = {{ {1, 0, 0}, {0, 1, 0}, {0, 0, 1} }};
const pl_transform3x3 pl_transform3x3_identity = { .mat = {{ {1, 0, 0}, {0, 1, 0}, {0, 0, 1} }} };

void pl_matrix3x3_apply(const pl_matrix3x3 *mat, float vec[3])
{
    const float x = vec[0], y = vec[1], z = vec[2];
    for (int i = 0; i < 3; i++)
        vec[i] = mat->m[i][0] * x + mat->m[i][1] * y + mat->m[i][2] * z;
}

void pl_matrix3x3_scale(pl_matrix3x3 *mat, float scale)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            mat->m[i][j] *= scale;
}

// adjugate / determinant in double precision
void pl_matrix3x3_invert(pl_matrix3x3 *mat)
{
    double m[3][3], adj[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            m[i][j] = mat->m[i][j];
    adj[0][0] =  (m[1][1] * m[2][2] - m[2][1] * m[1][2]);
    adj[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]);
    adj[0][2] =  (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    adj[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]);
    adj[1][1] =  (m[0][0] * m[2][2] - m[2][0] * m[0][2]);
    adj[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]);
    adj[2][0] =  (m[1][0] * m[2][1] - m[2][0] * m[1][1]);
    adj[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]);
    adj[2][2] =  (m[0][0] * m[1][1] - m[1][0] * m[0][1]);
    double det = m[0][0] * adj[0][0] + m[1][0] * adj[0][1] + m[2][0] * adj[0][2];
    det = 1.0 / det;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            mat->m[i][j] = det * adj[i][j];
}

void pl_matrix3x3_mul(pl_matrix3x3 *a, const pl_matrix3x3 *b)
{
    const pl_matrix3x3 o = *a;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            a->m[r][c] = o.m[r][0] * b->m[0][c] + o.m[r][1] * b->m[1][c] + o.m[r][2] * b->m[2][c];
}

void pl_matrix3x3_rmul(const pl_matrix3x3 *a, pl_matrix3x3 *b)
{
    pl_matrix3x3 m = *a;
    pl_matrix3x3_mul(&m, b);
    *b = m;
}

void pl_transform3x3_apply(const pl_transform3x3 *t, float vec[3])
{
    pl_matrix3x3_apply(&t->mat, vec);
    for (int i = 0; i < 3; i++)
        vec[i] += t->c[i];
}

void pl_transform3x3_scale(pl_transform3x3 *t, float scale)
{
    pl_matrix3x3_scale(&t->mat, scale);
    for (int i = 0; i < 3; i++)
        t->c[i] *= scale;
}

void pl_transform3x3_invert(pl_transform3x3 *t)
{
    pl_matrix3x3_invert(&t->mat);
    const float c0 = t->c[0], c1 = t->c[1], c2 = t->c[2];
    for (int i = 0; i < 3; i++)
        t->c[i] = -(t->mat.m[i][0] * c0 + t->mat.m[i][1] * c1 + t->mat.m[i][2] * c2);
}

// ---- colour systems / representation ----------------------------------------------------------

// Properties of the colour systems as bit sets over enum pl_color_system
#define SYS_BIT(s) (1u << (s))
static const unsigned SYS_NOT_YCBCR = SYS_BIT(PL_COLOR_SYSTEM_UNKNOWN) | SYS_BIT(PL_COLOR_SYSTEM_RGB) |
                                      SYS_BIT(PL_COLOR_SYSTEM_XYZ);
static const unsigned SYS_NONLINEAR = SYS_BIT(PL_COLOR_SYSTEM_BT_2020_C) | SYS_BIT(PL_COLOR_SYSTEM_BT_2100_PQ) |
                                      SYS_BIT(PL_COLOR_SYSTEM_BT_2100_HLG) |
                                      SYS_BIT(PL_COLOR_SYSTEM_DOLBYVISION) | SYS_BIT(PL_COLOR_SYSTEM_XYZ);

static bool sys_in(unsigned set, enum pl_color_system sys)
{
    return (unsigned) sys < 32 && (set >> sys & 1u);
}

bool pl_color_system_is_ycbcr_like(enum pl_color_system sys)
{
    return !sys_in(SYS_NOT_YCBCR, sys);
}

bool pl_color_system_is_linear(enum pl_color_system sys)
{
    return !sys_in(SYS_NONLINEAR, sys);
}

bool pl_color_repr_equal(const struct pl_color_repr *a, const struct pl_color_repr *b)
{
    return a->sys == b->sys && a->levels == b->levels && a->alpha == b->alpha &&
           a->dovi == b->dovi && a->bits.sample_depth == b->bits.sample_depth &&
           a->bits.color_depth == b->bits.color_depth && a->bits.bit_shift == b->bits.bit_shift;
}

enum pl_color_levels pl_color_levels_guess(const struct pl_color_repr *repr)
{
    const bool full = repr->sys == PL_COLOR_SYSTEM_DOLBYVISION ||
                      (!repr->levels && !pl_color_system_is_ycbcr_like(repr->sys));
    if (full)
        return PL_COLOR_LEVELS_FULL;
    return repr->levels ? repr->levels : PL_COLOR_LEVELS_LIMITED;
}

// Factor that brings sampled values to the full [0,1] range of the *colour* depth; afterwards
// the representation is "as if" sampled at full depth with no shift. Limited range is defined on
// the bit-shifted integer grid (ratio of powers of two), full range on the code maxima.
float pl_color_repr_normalize(struct pl_color_repr *repr)
{
    struct pl_bit_encoding *const be = &repr->bits;
    int sample = be->sample_depth ? be->sample_depth : be->color_depth;
    int color = be->color_depth ? be->color_depth : be->sample_depth;
    if (!sample)
        sample = color = 8;
    const long long smax = 1LL << sample, cmax = 1LL << color;
    float factor = 1.0;
    if (be->bit_shift)
        factor /= (1LL << be->bit_shift);
    const bool limited = pl_color_levels_guess(repr) == PL_COLOR_LEVELS_LIMITED;
    factor *= limited ? (float) smax / cmax : (smax - 1.) / (cmax - 1.);
    be->bit_shift = 0;
    be->color_depth = be->sample_depth;
    return factor;
}

// ---- transfer characteristics ------------------------------------------------------------------

static const float PQ_M1 = 2610./4096 * 1./4, PQ_M2 = 2523./4096 * 128, PQ_C1 = 3424./4096,
                   PQ_C2 = 2413./4096 * 32, PQ_C3 = 2392./4096 * 32;

// peak of the transfer function's signal range relative to SDR white
float pl_color_transfer_nominal_peak(enum pl_color_transfer trc)
{
    static const struct { enum pl_color_transfer trc; float peak; } peaks[] = {
        { PL_COLOR_TRC_PQ,     10000.0 / PL_COLOR_SDR_WHITE },
        { PL_COLOR_TRC_SCRGB,  10000.0 / PL_COLOR_SDR_WHITE },
        { PL_COLOR_TRC_HLG,    12.0 / 3.17955 },
        { PL_COLOR_TRC_V_LOG,  46.0855 },
        { PL_COLOR_TRC_S_LOG1, 6.52 },
        { PL_COLOR_TRC_S_LOG2, 9.212 },
    };
    for (size_t i = 0; i < PLB_ARRAY_SIZE(peaks); i++) {
        if (peaks[i].trc == trc)
            return peaks[i].peak;
    }
    return 1.0;
}

float pl_hdr_rescale(enum pl_hdr_scaling from, enum pl_hdr_scaling to, float x)
{
    if (from == to || !x)
        return x;
    x = fmaxf(x, 0.0f);

    // bring to PL_HDR_NORM
    if (from == PL_HDR_PQ) {
        x = powf(x, 1.0f / PQ_M2);
        x = fmaxf(x - PQ_C1, 0.0f) / (PQ_C2 - PQ_C3 * x);
        x = powf(x, 1.0f / PQ_M1);
        x *= 10000.0f;
        x /= PL_COLOR_SDR_WHITE;
    } else if (from == PL_HDR_NITS) {
        x /= PL_COLOR_SDR_WHITE;
    } else if (from == PL_HDR_SQRT) {
        x *= x;
    }

    switch (to) {
    case PL_HDR_SQRT: return sqrtf(x);
    case PL_HDR_NITS: return x * PL_COLOR_SDR_WHITE;
    case PL_HDR_PQ:
        x *= PL_COLOR_SDR_WHITE / 10000.0f;
        x = powf(x, PQ_M1);
        x = (PQ_C1 + PQ_C2 * x) / (1.0f + PQ_C3 * x);
        return powf(x, PQ_M2);
    default: return x;
    }
}

// ---- HDR metadata -------------------------------------------------------------------------------

const struct pl_hdr_metadata pl_hdr_metadata_empty = {0};
const struct pl_hdr_metadata pl_hdr_metadata_hdr10 = {
    .prim = { {0.708, 0.292}, {0.170, 0.797}, {0.131, 0.046}, {0.31271, 0.32902} },
    .min_luma = 0, .max_luma = 10000, .max_cll = 10000, .max_fall = 0,
};

bool pl_raw_primaries_equal(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b)
{
    return !memcmp(a, b, sizeof(*a)) ||
           (a->red.x == b->red.x && a->red.y == b->red.y && a->green.x == b->green.x &&
            a->green.y == b->green.y && a->blue.x == b->blue.x && a->blue.y == b->blue.y &&
            a->white.x == b->white.x && a->white.y == b->white.y);
}

bool pl_raw_primaries_similar(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b)
{
    const float *pa = (const float *) a, *pb = (const float *) b;
    float delta = 0;
    for (int i = 0; i < 8; i++)
        delta += fabsf(pa[i] - pb[i]);
    return delta < 0.001;
}

void pl_raw_primaries_merge(struct pl_raw_primaries *orig, const struct pl_raw_primaries *update)
{
    float *pa = (float *) orig;
    const float *pb = (const float *) update;
    for (int i = 0; i < 8; i++)
        pa[i] = PLB_DEF(pa[i], pb[i]);
}

static bool bezier_equal(const struct pl_hdr_bezier *a, const struct pl_hdr_bezier *b)
{
    return a->target_luma == b->target_luma && a->knee_x == b->knee_x &&
           a->knee_y == b->knee_y && a->num_anchors == b->num_anchors &&
           !memcmp(a->anchors, b->anchors, sizeof(a->anchors[0]) * a->num_anchors);
}

bool pl_hdr_metadata_equal(const struct pl_hdr_metadata *a, const struct pl_hdr_metadata *b)
{
    return pl_raw_primaries_equal(&a->prim, &b->prim) && a->min_luma == b->min_luma &&
           a->max_luma == b->max_luma && a->max_cll == b->max_cll && a->max_fall == b->max_fall &&
           a->scene_max[0] == b->scene_max[0] && a->scene_max[1] == b->scene_max[1] &&
           a->scene_max[2] == b->scene_max[2] && a->scene_avg == b->scene_avg &&
           bezier_equal(&a->ootf, &b->ootf) && a->max_pq_y == b->max_pq_y &&
           a->avg_pq_y == b->avg_pq_y;
}

void pl_hdr_metadata_merge(struct pl_hdr_metadata *orig, const struct pl_hdr_metadata *update)
{
    pl_raw_primaries_merge(&orig->prim, &update->prim);
    if (!orig->min_luma) orig->min_luma = update->min_luma;
    if (!orig->max_luma) orig->max_luma = update->max_luma;
    if (!orig->max_cll)  orig->max_cll = update->max_cll;
    if (!orig->max_fall) orig->max_fall = update->max_fall;
    if (!orig->scene_max[1])
        memcpy(orig->scene_max, update->scene_max, sizeof(orig->scene_max));
    if (!orig->scene_avg) orig->scene_avg = update->scene_avg;
    if (!orig->ootf.target_luma) orig->ootf = update->ootf;
    if (!orig->max_pq_y) orig->max_pq_y = update->max_pq_y;
    if (!orig->avg_pq_y) orig->avg_pq_y = update->avg_pq_y;
}

bool pl_hdr_metadata_contains(const struct pl_hdr_metadata *data, enum pl_hdr_metadata_type type)
{
    const bool hdr10 = data->max_luma;
    const bool hdr10plus = data->scene_avg &&
                           (data->scene_max[0] || data->scene_max[1] || data->scene_max[2]);
    const bool cie_y = data->max_pq_y && data->avg_pq_y;
    switch (type) {
    case PL_HDR_METADATA_NONE:      return true;
    case PL_HDR_METADATA_ANY:       return hdr10 || hdr10plus || cie_y;
    case PL_HDR_METADATA_HDR10:     return hdr10;
    case PL_HDR_METADATA_HDR10PLUS: return hdr10plus;
    case PL_HDR_METADATA_CIE_Y:     return cie_y;
    default: return false;
    }
}

// ---- colour spaces ------------------------------------------------------------------------------

const struct pl_color_space pl_color_space_unknown = {0};
const struct pl_color_space pl_color_space_srgb = { PL_COLOR_PRIM_BT_709, PL_COLOR_TRC_SRGB };
const struct pl_color_space pl_color_space_bt709 = { PL_COLOR_PRIM_BT_709, PL_COLOR_TRC_BT_1886 };
const struct pl_color_space pl_color_space_hdr10 = { PL_COLOR_PRIM_BT_2020, PL_COLOR_TRC_PQ };
const struct pl_color_space pl_color_space_bt2020_hlg = { PL_COLOR_PRIM_BT_2020, PL_COLOR_TRC_HLG };
const struct pl_color_space pl_color_space_monitor = { PL_COLOR_PRIM_BT_709, PL_COLOR_TRC_UNKNOWN };
const struct pl_color_adjustment pl_color_adjustment_neutral = { PL_COLOR_ADJUSTMENT_NEUTRAL };

bool pl_color_space_is_hdr(const struct pl_color_space *csp)
{
    return csp->hdr.max_luma > PL_COLOR_SDR_WHITE || pl_color_transfer_is_hdr(csp->transfer);
}

// Transfer functions whose black point is part of the curve itself (absolute or display-referred
// definitions); all others are scaled to the display's black level.
bool pl_color_space_is_black_scaled(const struct pl_color_space *csp)
{
    static const enum pl_color_transfer absolute_black[] = {
        PL_COLOR_TRC_BT_1886, PL_COLOR_TRC_PQ, PL_COLOR_TRC_SCRGB,
        PL_COLOR_TRC_V_LOG, PL_COLOR_TRC_S_LOG1, PL_COLOR_TRC_S_LOG2,
    };
    for (size_t i = 0; i < PLB_ARRAY_SIZE(absolute_black); i++) {
        if (absolute_black[i] == csp->transfer)
            return false;
    }
    return true;
}

void pl_color_space_merge(struct pl_color_space *orig, const struct pl_color_space *update)
{
    if (!orig->primaries) orig->primaries = update->primaries;
    if (!orig->transfer)  orig->transfer = update->transfer;
    pl_hdr_metadata_merge(&orig->hdr, &update->hdr);
}

bool pl_color_space_equal(const struct pl_color_space *a, const struct pl_color_space *b)
{
    return a->primaries == b->primaries && a->transfer == b->transfer &&
           pl_hdr_metadata_equal(&a->hdr, &b->hdr);
}

// the narrow (BT.709-class) gamuts; everything else counts as wide when a default is inferred
static bool narrow_gamut(enum pl_color_primaries prim)
{
    static const enum pl_color_primaries narrow[] = {
        PL_COLOR_PRIM_UNKNOWN, PL_COLOR_PRIM_BT_601_525, PL_COLOR_PRIM_BT_601_625,
        PL_COLOR_PRIM_BT_709, PL_COLOR_PRIM_BT_470M, PL_COLOR_PRIM_EBU_3213,
    };
    for (size_t i = 0; i < PLB_ARRAY_SIZE(narrow); i++) {
        if (narrow[i] == prim)
            return true;
    }
    return false;
}

static float clampf(float x, float lo, float hi) { return x < lo ? lo : x > hi ? hi : x; }

void pl_color_space_nominal_luma_ex(const struct pl_nominal_luma_params *params)
{
    if (!params || (!params->out_min && !params->out_max && !params->out_avg))
        return;
    const struct pl_color_space *csp = params->color;
    const enum pl_hdr_scaling sc = params->scaling;
    const enum pl_hdr_metadata_type md = params->metadata;
    float lo = 0, hi = 0, avg = 0;

    if (md != PL_HDR_METADATA_NONE) {
        lo = pl_hdr_rescale(PL_HDR_NITS, sc, csp->hdr.min_luma);
        hi = pl_hdr_rescale(PL_HDR_NITS, sc, csp->hdr.max_luma);
        if (!hi && csp->hdr.max_cll)
            hi = pl_hdr_rescale(PL_HDR_NITS, sc, csp->hdr.max_cll);
    }

    // HDR10+ scene statistics (maxRGB based)
    if ((md == PL_HDR_METADATA_ANY || md == PL_HDR_METADATA_HDR10PLUS) &&
        pl_hdr_metadata_contains(&csp->hdr, PL_HDR_METADATA_HDR10PLUS))
    {
        const float *sm = csp->hdr.scene_max;
        const float maxscl = PLB_MAX(PLB_MAX(sm[0], sm[1]), sm[2]);
        if (maxscl) {
            struct pl_raw_primaries prim = csp->hdr.prim;
            pl_raw_primaries_merge(&prim, pl_raw_primaries_get(csp->primaries));
            const pl_matrix3x3 rgb2xyz = pl_get_rgb2xyz_matrix(&prim);
            const float max_luma = rgb2xyz.m[1][0] * sm[0] + rgb2xyz.m[1][1] * sm[1] +
                                   rgb2xyz.m[1][2] * sm[2];
            const float coef = max_luma / maxscl;
            hi = pl_hdr_rescale(PL_HDR_NITS, sc, max_luma);
            avg = pl_hdr_rescale(PL_HDR_NITS, sc, coef * csp->hdr.scene_avg);
        }
    }

    // measured (peak-detected) luminance wins
    if ((md == PL_HDR_METADATA_ANY || md == PL_HDR_METADATA_CIE_Y) &&
        pl_hdr_metadata_contains(&csp->hdr, PL_HDR_METADATA_CIE_Y))
    {
        hi = pl_hdr_rescale(PL_HDR_PQ, sc, csp->hdr.max_pq_y);
        avg = pl_hdr_rescale(PL_HDR_PQ, sc, csp->hdr.avg_pq_y);
    }

    const float hdr_min = pl_hdr_rescale(PL_HDR_NITS, sc, PL_COLOR_HDR_BLACK);
    const float hdr_max = pl_hdr_rescale(PL_HDR_PQ, sc, 1.0f);
    hi = hi ? clampf(hi, hdr_min, hdr_max) : 0;
    lo = lo ? clampf(lo, hdr_min, hdr_max) : 0;
    if ((hi && lo >= hi) || lo >= hdr_max)
        lo = hi = 0;

    if (!hi) {
        if (csp->transfer == PL_COLOR_TRC_HLG) {
            hi = pl_hdr_rescale(PL_HDR_NITS, sc, PL_COLOR_HLG_PEAK);
        } else {
            const float peak = pl_color_transfer_nominal_peak(csp->transfer);
            hi = pl_hdr_rescale(PL_HDR_NORM, sc, peak);
        }
    }
    if (!lo) {
        if (pl_color_transfer_is_hdr(csp->transfer)) {
            lo = hdr_min;
        } else {
            // SDR: fixed 1000:1 contrast below the peak
            const float peak = pl_hdr_rescale(sc, PL_HDR_NITS, hi);
            lo = pl_hdr_rescale(PL_HDR_NITS, sc, peak / PL_COLOR_SDR_CONTRAST);
        }
    }
    if (avg)
        avg = clampf(avg, lo, hi);
    if (params->out_min) *params->out_min = lo;
    if (params->out_max) *params->out_max = hi;
    if (params->out_avg) *params->out_avg = avg;
}

// Defaults for an under-specified colour space: BT.709 / BT.1886, luminance range from the
// HDR10 metadata or the transfer function, mastering primaries from the nominal ones.
void pl_color_space_infer(struct pl_color_space *space)
{
    space->primaries = space->primaries ? space->primaries : PL_COLOR_PRIM_BT_709;
    space->transfer = space->transfer ? space->transfer : PL_COLOR_TRC_BT_1886;
    struct pl_nominal_luma_params lp = {
        .color = space,
        .metadata = PL_HDR_METADATA_HDR10,
        .scaling = PL_HDR_NITS,
        .out_max = &space->hdr.max_luma,
    };
    if (!space->hdr.min_luma)
        lp.out_min = &space->hdr.min_luma;
    pl_color_space_nominal_luma_ex(&lp);
    if (!pl_primaries_valid(&space->hdr.prim))
        space->hdr.prim = *pl_raw_primaries_get(space->primaries);
}

// The transfer function an unspecified space takes next to a reference: SDR curves carry
// over, HDR / camera-log references pair with BT.1886, ProPhoto with sRGB, the rest gamma 2.2
static enum pl_color_transfer companion_transfer(enum pl_color_transfer ref)
{
    static const struct { enum pl_color_transfer ref, companion; } map[] = {
        { PL_COLOR_TRC_BT_1886,   PL_COLOR_TRC_BT_1886 },
        { PL_COLOR_TRC_SRGB,      PL_COLOR_TRC_SRGB },
        { PL_COLOR_TRC_GAMMA22,   PL_COLOR_TRC_GAMMA22 },
        { PL_COLOR_TRC_PQ,        PL_COLOR_TRC_BT_1886 },
        { PL_COLOR_TRC_HLG,       PL_COLOR_TRC_BT_1886 },
        { PL_COLOR_TRC_V_LOG,     PL_COLOR_TRC_BT_1886 },
        { PL_COLOR_TRC_S_LOG1,    PL_COLOR_TRC_BT_1886 },
        { PL_COLOR_TRC_S_LOG2,    PL_COLOR_TRC_BT_1886 },
        { PL_COLOR_TRC_PRO_PHOTO, PL_COLOR_TRC_SRGB },
    };
    for (size_t i = 0; i < PLB_ARRAY_SIZE(map); i++) {
        if (map[i].ref == ref)
            return map[i].companion;
    }
    return PL_COLOR_TRC_GAMMA22;
}

static void infer_both(struct pl_color_space *space, struct pl_color_space *ref)
{
    pl_color_space_infer(ref);
    if (!space->primaries)
        space->primaries = narrow_gamut(ref->primaries) ? ref->primaries : PL_COLOR_PRIM_BT_709;
    if (!space->transfer)
        space->transfer = companion_transfer(ref->transfer);
    pl_color_space_infer(space);
}

void pl_color_space_infer_ref(struct pl_color_space *space, const struct pl_color_space *refp)
{
    struct pl_color_space ref = *refp;
    infer_both(space, &ref);
}

void pl_color_space_infer_map(struct pl_color_space *src, struct pl_color_space *dst)
{
    const bool unknown_src = !src->hdr.min_luma, unknown_dst = !dst->hdr.min_luma;
    infer_both(dst, src);
    // display-referred curves inherit the black point of the other side
    const bool dynamic_src = pl_color_space_is_black_scaled(src) ||
                             src->transfer == PL_COLOR_TRC_BT_1886;
    if (unknown_src && dynamic_src)
        src->hdr.min_luma = dst->hdr.min_luma;
    if (unknown_dst && !pl_color_space_is_hdr(src) && !pl_color_space_is_hdr(dst))
        dst->hdr.min_luma = src->hdr.min_luma;
    if (src->transfer == PL_COLOR_TRC_HLG && pl_color_space_is_hdr(dst))
        src->hdr.max_luma = dst->hdr.max_luma;
}

void pl_chroma_location_offset(enum pl_chroma_location loc, float *x, float *y)
{
    *x = *y = 0;
    loc = PLB_DEF(loc, PL_CHROMA_LEFT);
    if (loc == PL_CHROMA_LEFT || loc == PL_CHROMA_TOP_LEFT || loc == PL_CHROMA_BOTTOM_LEFT)
        *x = -0.5;
    if (loc == PL_CHROMA_TOP_LEFT || loc == PL_CHROMA_TOP_CENTER)
        *y = -0.5;
    if (loc == PL_CHROMA_BOTTOM_LEFT || loc == PL_CHROMA_BOTTOM_CENTER)
        *y = 0.5;
}

// ---- primaries ------------------------------------------------------------------------------------

#define D50 {0.3457, 0.3585}
#define D65 {0.3127, 0.3290}
#define ILL_C {0.3100, 0.3160}
#define ILL_E {1.0/3.0, 1.0/3.0}
#define DCI_W {0.3140, 0.3510}
#define ACES_W {0.32168, 0.33767}

const struct pl_raw_primaries *pl_raw_primaries_get(enum pl_color_primaries prim)
{
    // chromaticities from ITU-R BT.470-6 / BT.601-7 / BT.709-5 / BT.2020, SMPTE, ACES etc.
    static const struct pl_raw_primaries table[PL_COLOR_PRIM_COUNT] = {
        [PL_COLOR_PRIM_BT_601_525] = { {0.630, 0.340}, {0.310, 0.595}, {0.155, 0.070}, D65 },
        [PL_COLOR_PRIM_BT_601_625] = { {0.640, 0.330}, {0.290, 0.600}, {0.150, 0.060}, D65 },
        [PL_COLOR_PRIM_BT_709]     = { {0.640, 0.330}, {0.300, 0.600}, {0.150, 0.060}, D65 },
        [PL_COLOR_PRIM_BT_470M]    = { {0.670, 0.330}, {0.210, 0.710}, {0.140, 0.080}, ILL_C },
        [PL_COLOR_PRIM_EBU_3213]   = { {0.630, 0.340}, {0.295, 0.605}, {0.155, 0.077}, D65 },
        [PL_COLOR_PRIM_BT_2020]    = { {0.708, 0.292}, {0.170, 0.797}, {0.131, 0.046}, D65 },
        [PL_COLOR_PRIM_APPLE]      = { {0.625, 0.340}, {0.280, 0.595}, {0.115, 0.070}, D65 },
        [PL_COLOR_PRIM_ADOBE]      = { {0.640, 0.330}, {0.210, 0.710}, {0.150, 0.060}, D65 },
        [PL_COLOR_PRIM_PRO_PHOTO]  = { {0.7347, 0.2653}, {0.1596, 0.8404}, {0.0366, 0.0001}, D50 },
        [PL_COLOR_PRIM_CIE_1931]   = { {0.7347, 0.2653}, {0.2738, 0.7174}, {0.1666, 0.0089}, ILL_E },
        [PL_COLOR_PRIM_DCI_P3]     = { {0.680, 0.320}, {0.265, 0.690}, {0.150, 0.060}, DCI_W },
        [PL_COLOR_PRIM_DISPLAY_P3] = { {0.680, 0.320}, {0.265, 0.690}, {0.150, 0.060}, D65 },
        [PL_COLOR_PRIM_V_GAMUT]    = { {0.730, 0.280}, {0.165, 0.840}, {0.100, -0.03}, D65 },
        [PL_COLOR_PRIM_S_GAMUT]    = { {0.730, 0.280}, {0.140, 0.855}, {0.100, -0.05}, D65 },
        [PL_COLOR_PRIM_FILM_C]     = { {0.681, 0.319}, {0.243, 0.692}, {0.145, 0.049}, ILL_C },
        [PL_COLOR_PRIM_ACES_AP0]   = { {0.7347, 0.2653}, {0.0000, 1.0000}, {0.0001, -0.0770}, ACES_W },
        [PL_COLOR_PRIM_ACES_AP1]   = { {0.713, 0.293}, {0.165, 0.830}, {0.128, 0.044}, ACES_W },
    };
    if (!prim || prim >= PL_COLOR_PRIM_COUNT)
        prim = PL_COLOR_PRIM_BT_709;
    return &table[prim];
}

static inline float cie_X(struct pl_cie_xy xy) { return xy.x / xy.y; }
static inline float cie_Z(struct pl_cie_xy xy) { return (1 - xy.x - xy.y) / xy.y; }

// Columns are the XYZ of the primaries scaled so that RGB (1,1,1) maps to the white point
pl_matrix3x3 pl_get_rgb2xyz_matrix(const struct pl_raw_primaries *prim)
{
    const struct pl_cie_xy pts[4] = { prim->red, prim->green, prim->blue, prim->white };
    float X[4], Z[4], S[3];
    for (int i = 0; i < 4; i++)
        X[i] = cie_X(pts[i]);
    for (int i = 0; i < 4; i++)
        Z[i] = cie_Z(pts[i]);
    pl_matrix3x3 out = {{{0}}};
    for (int i = 0; i < 3; i++) {
        out.m[0][i] = X[i];
        out.m[1][i] = 1;
        out.m[2][i] = Z[i];
    }
    pl_matrix3x3_invert(&out);
    for (int i = 0; i < 3; i++)
        S[i] = out.m[i][0] * X[3] + out.m[i][1] * 1 + out.m[i][2] * Z[3];
    for (int i = 0; i < 3; i++) {
        out.m[0][i] = S[i] * X[i];
        out.m[1][i] = S[i] * 1;
        out.m[2][i] = S[i] * Z[i];
    }
    return out;
}

pl_matrix3x3 pl_get_xyz2rgb_matrix(const struct pl_raw_primaries *prim)
{
    pl_matrix3x3 out = pl_get_rgb2xyz_matrix(prim);
    pl_matrix3x3_invert(&out);
    return out;
}

// CAT16 cone response (Li et al. 2017)
static const pl_matrix3x3 cat16 = {{
    {  0.401288, 0.650173, -0.051461 },
    { -0.250268, 1.204414,  0.045854 },
    { -0.002079, 0.048952,  0.953127 },
}};

// mat := mat * CAT16^-1 * diag(dst/src cone response) * CAT16
static void chromatic_adaptation(struct pl_cie_xy src, struct pl_cie_xy dst, pl_matrix3x3 *mat)
{
    if (fabs(src.x - dst.x) < 1e-6 && fabs(src.y - dst.y) < 1e-6)
        return;
    float C[3][2];
    for (int i = 0; i < 3; i++) {
        C[i][0] = cat16.m[i][0] * cie_X(src) + cat16.m[i][1] * 1 + cat16.m[i][2] * cie_Z(src);
        C[i][1] = cat16.m[i][0] * cie_X(dst) + cat16.m[i][1] * 1 + cat16.m[i][2] * cie_Z(dst);
    }
    pl_matrix3x3 tmp = {0};
    for (int i = 0; i < 3; i++)
        tmp.m[i][i] = C[i][1] / C[i][0];
    pl_matrix3x3_mul(&tmp, &cat16);
    pl_matrix3x3 inv = cat16;
    pl_matrix3x3_invert(&inv);
    pl_matrix3x3_mul(mat, &inv);
    pl_matrix3x3_mul(mat, &tmp);
}

pl_matrix3x3 pl_get_adaptation_matrix(struct pl_cie_xy src, struct pl_cie_xy dst)
{
    struct pl_raw_primaries csp = *pl_raw_primaries_get(PL_COLOR_PRIM_BT_709);
    csp.white = src;
    pl_matrix3x3 rgb2xyz = pl_get_rgb2xyz_matrix(&csp);
    pl_matrix3x3 xyz2rgb = rgb2xyz;
    pl_matrix3x3_invert(&xyz2rgb);
    chromatic_adaptation(src, dst, &xyz2rgb);
    pl_matrix3x3_mul(&xyz2rgb, &rgb2xyz);
    return xyz2rgb;
}

pl_matrix3x3 pl_get_color_mapping_matrix(const struct pl_raw_primaries *src,
                                         const struct pl_raw_primaries *dst,
                                         enum pl_rendering_intent intent)
{
    if (intent == PL_INTENT_SATURATION)
        return pl_matrix3x3_identity;
    pl_matrix3x3 out = pl_get_xyz2rgb_matrix(dst);
    if (intent != PL_INTENT_ABSOLUTE_COLORIMETRIC)
        chromatic_adaptation(src->white, dst->white, &out);
    const pl_matrix3x3 rgb2xyz = pl_get_rgb2xyz_matrix(src);
    pl_matrix3x3_mul(&out, &rgb2xyz);
    return out;
}

// IPT (Ebner & Fairchild 1998): Hunt-Pointer-Estevez LMS with 4% crosstalk, D65-adapted
// RGB -> LMS of IPT-PQ-C2: Hunt-Pointer-Estevez cone fundamentals with 4 % crosstalk between the
// cone responses, adapted from the primaries' white point to D65.
pl_matrix3x3 pl_ipt_rgb2lms(const struct pl_raw_primaries *prim)
{
    static const pl_matrix3x3 hpe = {{
        {  0.40024f, 0.70760f, -0.08081f },
        { -0.22630f, 1.16532f,  0.04570f },
        {  0.00000f, 0.00000f,  0.91822f },
    }};
    const float crosstalk = 0.04;
    pl_matrix3x3 lms;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            lms.m[i][j] = i == j ? 1 - 2*crosstalk : crosstalk;
    }
    pl_matrix3x3_mul(&lms, &hpe);
    const struct pl_cie_xy d65 = D65;
    chromatic_adaptation(prim->white, d65, &lms);
    const pl_matrix3x3 to_xyz = pl_get_rgb2xyz_matrix(prim);
    pl_matrix3x3_mul(&lms, &to_xyz);
    return lms;
}

pl_matrix3x3 pl_ipt_lms2rgb(const struct pl_raw_primaries *prim)
{
    pl_matrix3x3 m = pl_ipt_rgb2lms(prim);
    pl_matrix3x3_invert(&m);
    return m;
}

const pl_matrix3x3 pl_ipt_lms2ipt = {{
    { 0.4000,  0.4000,  0.2000 },
    { 4.4550, -4.8510,  0.3960 },
    { 0.8056,  0.3572, -1.1628 },
}};

const pl_matrix3x3 pl_ipt_ipt2lms = {{
    { 1.0,  0.0975689,  0.205226 },
    { 1.0, -0.1138760,  0.133217 },
    { 1.0,  0.0326151, -0.676887 },
}};

// signed area test of p against the directed line a->b
static float side(struct pl_cie_xy p, struct pl_cie_xy a, struct pl_cie_xy b)
{
    return (p.x - b.x) * (a.y - b.y) - (a.x - b.x) * (p.y - b.y);
}

static bool in_gamut(struct pl_cie_xy p, const struct pl_raw_primaries *prim)
{
    const float d1 = side(p, prim->red, prim->green), d2 = side(p, prim->green, prim->blue),
                d3 = side(p, prim->blue, prim->red);
    const bool neg = d1 < -1e-6f || d2 < -1e-6f || d3 < -1e-6f;
    const bool pos = d1 > 1e-6f || d2 > 1e-6f || d3 > 1e-6f;
    return !(neg && pos);
}

bool pl_primaries_superset(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b)
{
    return in_gamut(b->red, a) && in_gamut(b->green, a) && in_gamut(b->blue, a);
}

bool pl_primaries_valid(const struct pl_raw_primaries *prim)
{
    const float area = (prim->blue.x - prim->green.x) * (prim->red.y - prim->green.y) -
                       (prim->red.x - prim->green.x) * (prim->blue.y - prim->green.y);
    return fabs(area) > 1e-6 && in_gamut(prim->white, prim);
}

static float dist2(struct pl_cie_xy a, struct pl_cie_xy b)
{
    const float dx = a.x - b.x, dy = a.y - b.y;
    return dx * dx + dy * dy;
}

// each primary of `a` is closest to the same-named primary of `b`
bool pl_primaries_compatible(const struct pl_raw_primaries *a, const struct pl_raw_primaries *b)
{
    const struct pl_cie_xy pa[3] = { a->red, a->green, a->blue };
    const struct pl_cie_xy pb[3] = { b->red, b->green, b->blue };
    for (int i = 0; i < 3; i++) {
        const float own = dist2(pa[i], pb[i]);
        for (int j = 0; j < 3; j++) {
            if (j != i && !(own < dist2(pa[i], pb[j])))
                return false;
        }
    }
    return true;
}

static struct pl_cie_xy intersect(struct pl_cie_xy a, struct pl_cie_xy b, struct pl_cie_xy c,
                                  struct pl_cie_xy d)
{
    const float det = (a.x - b.x) * (c.y - d.y) - (a.y - b.y) * (c.x - d.x);
    const float t = ((a.x - c.x) * (c.y - d.y) - (a.y - c.y) * (c.x - d.x)) / det;
    return (struct pl_cie_xy) {
        .x = t ? a.x + t * (b.x - a.x) : 0.0f,
        .y = t ? a.y + t * (b.y - a.y) : 0.0f,
    };
}

// clip vertex y of triangle (x, y, z) against the corner b of triangle (a, b, c)
static struct pl_cie_xy clip_vertex(struct pl_cie_xy x, struct pl_cie_xy y, struct pl_cie_xy z,
                                    struct pl_cie_xy a, struct pl_cie_xy b, struct pl_cie_xy c)
{
    const float d1 = side(y, a, b), d2 = side(y, b, c);
    if (d1 <= 0.0f && d2 <= 0.0f)
        return y;
    if (d1 > 0.0f && d2 > 0.0f)
        return b;
    return d1 > 0.0f ? intersect(a, b, y, z) : intersect(x, y, b, c);
}

struct pl_raw_primaries pl_primaries_clip(const struct pl_raw_primaries *s,
                                          const struct pl_raw_primaries *d)
{
    return (struct pl_raw_primaries) {
        .red   = clip_vertex(s->green, s->red, s->blue, d->green, d->red, d->blue),
        .green = clip_vertex(s->blue, s->green, s->red, d->blue, d->green, d->red),
        .blue  = clip_vertex(s->red, s->blue, s->green, d->red, d->blue, d->green),
        .white = s->white,
    };
}

// ---- YCbCr decode matrix -----------------------------------------------------------------------

// yuv->rgb matrix for luma weights (lr, lg, lb): Y -> (1,1,1), Cb/Cr orthogonal to luma,
// scaled so that Cb,Cr in [-0.5,0.5] just cover the RGB cube
static pl_matrix3x3 from_luma_weights(float lr, float lg, float lb)
{
    return (pl_matrix3x3) {{
        {1, 0,                    2 * (1-lr)          },
        {1, -2 * (1-lb) * lb/lg, -2 * (1-lr) * lr/lg  },
        {1,  2 * (1-lb),          0                   },
    }};
}

static void rotate_chroma(pl_matrix3x3 *m, const struct pl_color_adjustment *adj)
{
    const float hc = adj->saturation * cos(adj->hue), hs = adj->saturation * sin(adj->hue);
    for (int i = 0; i < 3; i++) {
        const float u = m->m[i][1], v = m->m[i][2];
        m->m[i][1] = hc * u - hs * v;
        m->m[i][2] = hs * u + hc * v;
    }
}

static struct pl_cie_xy daylight_from_temp(float temp)
{
    temp = clampf(temp, 1000, 25000);
    const double ti = 1000.0 / temp, ti2 = ti * ti, ti3 = ti2 * ti;
    const double x = temp <= 7000 ? -4.6070 * ti3 + 2.9678 * ti2 + 0.09911 * ti + 0.244063
                                  : -2.0064 * ti3 + 1.9018 * ti2 + 0.24748 * ti + 0.237040;
    return (struct pl_cie_xy) { .x = x, .y = -3 * (x*x) + 2.87 * x - 0.275 };
}

static struct pl_cie_xy blackbody_from_temp(float temp)
{
    temp = clampf(temp, 1667, 25000);
    const double ti = 1000.0 / temp, ti2 = ti * ti, ti3 = ti2 * ti;
    const double x = temp <= 4000 ? -0.2661239 * ti3 - 0.2343580 * ti2 + 0.8776956 * ti + 0.179910
                                  : -3.0258469 * ti3 + 2.1070379 * ti2 + 0.2226347 * ti + 0.240390;
    const double x2 = x * x, x3 = x2 * x;
    double y;
    if (temp <= 2222)
        y = -1.1063814 * x3 - 1.34811020 * x2 + 2.18555832 * x - 0.20219683;
    else if (temp <= 4000)
        y = -0.9549476 * x3 - 1.37418593 * x2 + 2.09137015 * x - 0.16748867;
    else
        y = 3.0817580 * x3 - 5.87338670 * x2 + 3.75112997 * x - 0.37001483;
    return (struct pl_cie_xy) { x, y };
}

static struct pl_cie_xy white_from_temp(float temp)
{
    const struct pl_cie_xy a = blackbody_from_temp(temp), b = daylight_from_temp(temp);
    float f = (temp - 2500) / (4000 - 2500);
    f = clampf(f, 0.0f, 1.0f);
    return (struct pl_cie_xy) { (1 - f) * a.x + f * b.x, (1 - f) * a.y + f * b.y };
}

pl_transform3x3 pl_color_repr_decode(struct pl_color_repr *repr,
                                     const struct pl_color_adjustment *adj)
{
    adj = adj ? adj : &pl_color_adjustment_neutral;
    pl_matrix3x3 m;
    switch (repr->sys) {
    case PL_COLOR_SYSTEM_BT_709:     m = from_luma_weights(0.2126, 0.7152, 0.0722); break;
    case PL_COLOR_SYSTEM_BT_601:     m = from_luma_weights(0.2990, 0.5870, 0.1140); break;
    case PL_COLOR_SYSTEM_SMPTE_240M: m = from_luma_weights(0.2122, 0.7013, 0.0865); break;
    case PL_COLOR_SYSTEM_BT_2020_NC: m = from_luma_weights(0.2627, 0.6780, 0.0593); break;
    case PL_COLOR_SYSTEM_BT_2020_C:
        m = (pl_matrix3x3) {{ {0, 0, 1}, {1, 0, 0}, {0, 1, 0} }};
        break;
    case PL_COLOR_SYSTEM_BT_2100_PQ: {
        static const float t = 0.008609, p = 0.111029625;
        m = (pl_matrix3x3) {{ {1.0, t, p}, {1.0, -t, -p}, {1.0, 0.560031, -0.320627} }};
        break;
    }
    case PL_COLOR_SYSTEM_BT_2100_HLG: {
        static const float t = 0.01571858011, p = 0.2095810681;
        m = (pl_matrix3x3) {{ {1.0, t, p}, {1.0, -t, -p}, {1.0, 1.02127108, -0.605274491} }};
        break;
    }
    case PL_COLOR_SYSTEM_DOLBYVISION:
        m = repr->dovi->nonlinear;
        break;
    case PL_COLOR_SYSTEM_YCGCO:
        m = (pl_matrix3x3) {{ {1, -1, 1}, {1, 1, 0}, {1, -1, -1} }};
        break;
    case PL_COLOR_SYSTEM_YCGCO_RE:
    case PL_COLOR_SYSTEM_YCGCO_RO:
        m = (pl_matrix3x3) {{ {1, -0.5, 0.5}, {1, 0.5, 0}, {1, -0.5, -0.5} }};
        break;
    case PL_COLOR_SYSTEM_XYZ:
        m = pl_get_xyz2rgb_matrix(pl_raw_primaries_get(PL_COLOR_PRIM_DCI_P3));
        break;
    default:
        m = pl_matrix3x3_identity;
        break;
    }

    if (pl_color_system_is_ycbcr_like(repr->sys)) {
        rotate_chroma(&m, adj);
    } else if (adj->saturation != 1.0 || adj->hue != 0.0) {
        // RGB-like input: go through BT.709 YCbCr for the hue/saturation rotation
        pl_matrix3x3 yuv2rgb = from_luma_weights(0.2126, 0.7152, 0.0722);
        pl_matrix3x3 rgb2yuv = yuv2rgb;
        pl_matrix3x3_invert(&rgb2yuv);
        rotate_chroma(&yuv2rgb, adj);
        pl_matrix3x3_rmul(&rgb2yuv, &m);
        pl_matrix3x3_rmul(&yuv2rgb, &m);
    }

    if (adj->temperature) {
        const struct pl_cie_xy src = white_from_temp(6500);
        const struct pl_cie_xy dst = white_from_temp(6500 + 3500 * adj->temperature);
        const pl_matrix3x3 adapt = pl_get_adaptation_matrix(src, dst);
        pl_matrix3x3_rmul(&adapt, &m);
    }

    pl_transform3x3 out = { .mat = m };
    const int depth = PLB_DEF(repr->bits.sample_depth, PLB_DEF(repr->bits.color_depth, 8));
    // code value 2^n maps to 1.0 in the limited-range definitions, textures map 2^n - 1
    const double scale = (1LL << depth) / ((1LL << depth) - 1.0);
    double ymax = 1.0, ymin = 0.0, cmax = 1.0, cmid = 128 / 256. * scale;
    if (pl_color_levels_guess(repr) == PL_COLOR_LEVELS_LIMITED) {
        ymax = 235 / 256. * scale;
        ymin = 16 / 256. * scale;
        cmax = 240 / 256. * scale;
    }
    double ymul = 1.0 / (ymax - ymin), cmul = 0.5 / (cmax - cmid);

    if (repr->sys == PL_COLOR_SYSTEM_YCGCO_RE || repr->sys == PL_COLOR_SYSTEM_YCGCO_RO) {
        const int extra = repr->sys == PL_COLOR_SYSTEM_YCGCO_RE ? 2 : 1;
        const double max_y = (1LL << (depth - extra)) - 1, max_c = (1LL << depth) - 1;
        ymul = cmul = max_c / max_y;
        ymin = 0;
        cmid = (1 << (depth - 1)) / max_c;
    }

    double mul[3] = { ymul, ymul, ymul }, black[3] = { ymin, ymin, ymin };
    if (repr->sys == PL_COLOR_SYSTEM_DOLBYVISION) {
        for (int i = 0; i < 3; i++) {
            mul[i] = 1.0;
            black[i] = repr->dovi->nonlinear_offset[i] * scale;
        }
    } else if (pl_color_system_is_ycbcr_like(repr->sys)) {
        mul[1] = mul[2] = cmul;
        black[1] = black[2] = cmid;
    }

    for (int i = 0; i < 3; i++) {
        mul[i] *= adj->contrast;
        out.c[i] += adj->brightness;
    }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            out.mat.m[i][j] *= mul[j];
            out.c[i] -= out.mat.m[i][j] * black[j];
        }
    }

    pl_matrix3x3_scale(&out.mat, pl_color_repr_normalize(repr));
    repr->sys = PL_COLOR_SYSTEM_RGB;
    repr->levels = PL_COLOR_LEVELS_FULL;
    return out;
}
