// Crop / aspect helpers of libplacebo/common.h that callers use to fill pl_frame.crop before
// pl_render_image (mpv, vf_libplacebo: fit a source rect into a target rect with pan & scan).
// Restated from the behaviour of src/common.c:35-100, 395-500 (same float / double arithmetic,
// checked bit for bit against the compiled reference in tests/test_t0_parity.py and by the
// reference's own src/tests/common.c through tests/test_reference_unit_tests.py).
#include <math.h>

#include "../internal.h"

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static float fminv(float a, float b) { return a < b ? a : b; }   // PL_MIN semantics (no NaN rules)
static float fmaxv(float a, float b) { return a > b ? a : b; }

// order every axis so that the "0" corner is the smaller one
void pl_rect2d_normalize(pl_rect2d *rc)
{
    const pl_rect2d in = *rc;
    rc->x0 = imin(in.x0, in.x1); rc->x1 = imax(in.x0, in.x1);
    rc->y0 = imin(in.y0, in.y1); rc->y1 = imax(in.y0, in.y1);
}

void pl_rect3d_normalize(pl_rect3d *rc)
{
    const pl_rect3d in = *rc;
    rc->x0 = imin(in.x0, in.x1); rc->x1 = imax(in.x0, in.x1);
    rc->y0 = imin(in.y0, in.y1); rc->y1 = imax(in.y0, in.y1);
    rc->z0 = imin(in.z0, in.z1); rc->z1 = imax(in.z0, in.z1);
}

void pl_rect2df_normalize(pl_rect2df *rc)
{
    const pl_rect2df in = *rc;
    rc->x0 = fminv(in.x0, in.x1); rc->x1 = fmaxv(in.x0, in.x1);
    rc->y0 = fminv(in.y0, in.y1); rc->y1 = fmaxv(in.y0, in.y1);
}

void pl_rect3df_normalize(pl_rect3df *rc)
{
    const pl_rect3df in = *rc;
    rc->x0 = fminv(in.x0, in.x1); rc->x1 = fmaxv(in.x0, in.x1);
    rc->y0 = fminv(in.y0, in.y1); rc->y1 = fmaxv(in.y0, in.y1);
    rc->z0 = fminv(in.z0, in.z1); rc->z1 = fmaxv(in.z0, in.z1);
}

pl_rect2d pl_rect2df_round(const pl_rect2df *rc)
{
    pl_rect2d out = { roundf(rc->x0), roundf(rc->y0), roundf(rc->x1), roundf(rc->y1) };
    return out;
}

pl_rect3d pl_rect3df_round(const pl_rect3df *rc)
{
    pl_rect3d out = { roundf(rc->x0), roundf(rc->y0), roundf(rc->z0),
                      roundf(rc->x1), roundf(rc->y1), roundf(rc->z1) };
    return out;
}

// |width| / |height|, 0 for a degenerate rect
float pl_rect2df_aspect(const pl_rect2df *rc)
{
    const float w = fabsf(rc->x1 - rc->x0), h = fabsf(rc->y1 - rc->y0);
    return h ? (w / h) : 0.0;
}

// scale about the centre
void pl_rect2df_stretch(pl_rect2df *rc, float stretch_x, float stretch_y)
{
    const float midx = (rc->x0 + rc->x1) / 2.0, midy = (rc->y0 + rc->y1) / 2.0;
    rc->x0 = rc->x0 * stretch_x + midx * (1.0 - stretch_x);
    rc->x1 = rc->x1 * stretch_x + midx * (1.0 - stretch_x);
    rc->y0 = rc->y0 * stretch_y + midy * (1.0 - stretch_y);
    rc->y1 = rc->y1 * stretch_y + midy * (1.0 - stretch_y);
}

// Change the aspect ratio keeping the centre: panscan 0 shrinks the overflowing dimension
// (letterbox), 1 grows the other one (crop), values between interpolate geometrically.
void pl_rect2df_aspect_set(pl_rect2df *rc, float aspect, float panscan)
{
    const float have = pl_rect2df_aspect(rc);
    if (!aspect || !have || aspect == have)
        return;
    float sx, sy;
    if (aspect > have) { // wider than now
        sx = powf(aspect / have, panscan);
        sy = powf(aspect / have, panscan - 1.0);
    } else {             // taller than now
        sx = powf(have / aspect, panscan - 1.0);
        sy = powf(have / aspect, panscan);
    }
    pl_rect2df_stretch(rc, sx, sy);
}

// Give `rc` the size and aspect of `src`, limited to what fits inside the original `rc`
void pl_rect2df_aspect_fit(pl_rect2df *rc, const pl_rect2df *src, float panscan)
{
    const float w = fabs(rc->x1 - rc->x0), h = fabs(rc->y1 - rc->y0);
    if (!w || !h)
        return;
    const float sx = fabs(src->x1 - src->x0) / w, sy = fabs(src->y1 - src->y0) / h;
    if (sx > 1.0 || sy > 1.0)
        pl_rect2df_aspect_set(rc, pl_rect2df_aspect(src), panscan); // too large: shrink to fit
    else
        pl_rect2df_stretch(rc, sx, sy);
}

// translate; a flipped axis moves the other way
void pl_rect2df_offset(pl_rect2df *rc, float offset_x, float offset_y)
{
    if (rc->x1 < rc->x0)
        offset_x = -offset_x;
    if (rc->y1 < rc->y0)
        offset_y = -offset_y;
    rc->x0 += offset_x; rc->x1 += offset_x;
    rc->y0 += offset_y; rc->y1 += offset_y;
}

// Rotate the coordinate system in quarter turns (clockwise); any integer is legal
void pl_rect2df_rotate(pl_rect2df *rc, pl_rotation rot)
{
    rot = (rot % 4 + 4) % 4;
    if (!rot)
        return;
    float x0 = rc->x0, y0 = rc->y0, x1 = rc->x1, y1 = rc->y1;
    if (rot >= 2) { // half turn: swap the corners
        float t;
        t = x0; x0 = x1; x1 = t;
        t = y0; y0 = y1; y1 = t;
        rot -= 2;
    }
    if (rot == 1) {
        const pl_rect2df quarter = { y1, x0, y0, x1 }; // x0, y0, x1, y1
        *rc = quarter;
    } else {
        const pl_rect2df same = { x0, y0, x1, y1 };
        *rc = same;
    }
}

// ---------------------------------------------------------------------------------------
// 2x2 matrices / affine transforms (common.h:64-116: sampling-rect rotation and offsets) and
// the rect forms of the 3x3 ones. Same operation order as src/common.c:115-128, 207-217,
// 238-393 (bit-exact vs the compiled reference, tests/test_t0_parity.py::test_matrix2x2_family).
// ---------------------------------------------------------------------------------------
const pl_matrix2x2 pl_matrix2x2_identity = {{ {1, 0}, {0, 1} }};
const pl_transform2x2 pl_transform2x2_identity = { .mat = {{ {1, 0}, {0, 1} }} };

pl_matrix2x2 pl_matrix2x2_rotation(float angle)
{
    const pl_matrix2x2 r = {{ { cosf(angle), -sinf(angle) }, { sinf(angle), cosf(angle) } }};
    return r;
}

void pl_matrix2x2_apply(const pl_matrix2x2 *mat, float vec[2])
{
    const float x = vec[0], y = vec[1];
    vec[0] = mat->m[0][0] * x + mat->m[0][1] * y;
    vec[1] = mat->m[1][0] * x + mat->m[1][1] * y;
}

// both corners of the rect are mapped as points
void pl_matrix2x2_apply_rc(const pl_matrix2x2 *mat, pl_rect2df *rc)
{
    float p0[2] = { rc->x0, rc->y0 }, p1[2] = { rc->x1, rc->y1 };
    pl_matrix2x2_apply(mat, p0);
    pl_matrix2x2_apply(mat, p1);
    rc->x0 = p0[0]; rc->y0 = p0[1];
    rc->x1 = p1[0]; rc->y1 = p1[1];
}

void pl_matrix2x2_mul(pl_matrix2x2 *a, const pl_matrix2x2 *b) // a := a * b
{
    const pl_matrix2x2 l = *a;
    for (int i = 0; i < 2; i++) {
        for (int j = 0; j < 2; j++)
            a->m[i][j] = l.m[i][0] * b->m[0][j] + l.m[i][1] * b->m[1][j];
    }
}

void pl_matrix2x2_rmul(const pl_matrix2x2 *a, pl_matrix2x2 *b) // b := a * b
{
    pl_matrix2x2 m = *a;
    pl_matrix2x2_mul(&m, b);
    *b = m;
}

void pl_matrix2x2_scale(pl_matrix2x2 *mat, float scale)
{
    for (int i = 0; i < 4; i++)
        (&mat->m[0][0])[i] *= scale;
}

void pl_matrix2x2_invert(pl_matrix2x2 *mat)
{
    const float a = mat->m[0][0], b = mat->m[0][1], c = mat->m[1][0], d = mat->m[1][1];
    const float invdet = 1.0f / (d * a - c * b);
    mat->m[0][0] =  d * invdet;
    mat->m[0][1] = -b * invdet;
    mat->m[1][0] = -c * invdet;
    mat->m[1][1] =  a * invdet;
}

void pl_transform2x2_apply(const pl_transform2x2 *t, float vec[2])
{
    pl_matrix2x2_apply(&t->mat, vec);
    vec[0] += t->c[0];
    vec[1] += t->c[1];
}

void pl_transform2x2_apply_rc(const pl_transform2x2 *t, pl_rect2df *rc)
{
    pl_matrix2x2_apply_rc(&t->mat, rc);
    rc->x0 += t->c[0]; rc->x1 += t->c[0];
    rc->y0 += t->c[1]; rc->y1 += t->c[1];
}

void pl_transform2x2_mul(pl_transform2x2 *a, const pl_transform2x2 *b) // a := a o b
{
    float c[2] = { b->c[0], b->c[1] };
    pl_transform2x2_apply(a, c);
    a->c[0] = c[0];
    a->c[1] = c[1];
    pl_matrix2x2_mul(&a->mat, &b->mat);
}

void pl_transform2x2_rmul(const pl_transform2x2 *a, pl_transform2x2 *b) // b := a o b
{
    pl_transform2x2_apply(a, b->c);
    pl_matrix2x2_rmul(&a->mat, &b->mat);
}

void pl_transform2x2_scale(pl_transform2x2 *t, float scale)
{
    pl_matrix2x2_scale(&t->mat, scale);
    t->c[0] *= scale;
    t->c[1] *= scale;
}

void pl_transform2x2_invert(pl_transform2x2 *t) // x = M^-1 y - M^-1 c
{
    pl_matrix2x2_invert(&t->mat);
    const float c0 = t->c[0], c1 = t->c[1];
    t->c[0] = -(t->mat.m[0][0] * c0 + t->mat.m[0][1] * c1);
    t->c[1] = -(t->mat.m[1][0] * c0 + t->mat.m[1][1] * c1);
}

// axis-aligned bounding box of the transformed rect (its four corners)
pl_rect2df pl_transform2x2_bounds(const pl_transform2x2 *t, const pl_rect2df *rc)
{
    float p[4][2] = { { rc->x0, rc->y0 }, { rc->x0, rc->y1 }, { rc->x1, rc->y0 }, { rc->x1, rc->y1 } };
    for (int i = 0; i < 4; i++)
        pl_transform2x2_apply(t, p[i]);
    pl_rect2df out;
    out.x0 = fminf(fminf(p[0][0], p[1][0]), fminf(p[2][0], p[3][0]));
    out.x1 = fmaxf(fmaxf(p[0][0], p[1][0]), fmaxf(p[2][0], p[3][0]));
    out.y0 = fminf(fminf(p[0][1], p[1][1]), fminf(p[2][1], p[3][1]));
    out.y1 = fmaxf(fmaxf(p[0][1], p[1][1]), fmaxf(p[2][1], p[3][1]));
    return out;
}

void pl_matrix3x3_apply_rc(const pl_matrix3x3 *mat, pl_rect3df *rc)
{
    float p0[3] = { rc->x0, rc->y0, rc->z0 }, p1[3] = { rc->x1, rc->y1, rc->z1 };
    pl_matrix3x3_apply(mat, p0);
    pl_matrix3x3_apply(mat, p1);
    rc->x0 = p0[0]; rc->y0 = p0[1]; rc->z0 = p0[2];
    rc->x1 = p1[0]; rc->y1 = p1[1]; rc->z1 = p1[2];
}

void pl_transform3x3_apply_rc(const pl_transform3x3 *t, pl_rect3df *rc)
{
    pl_matrix3x3_apply_rc(&t->mat, rc);
    rc->x0 += t->c[0]; rc->x1 += t->c[0];
    rc->y0 += t->c[1]; rc->y1 += t->c[1];
    rc->z0 += t->c[2]; rc->z1 += t->c[2];
}
