// Host-side filter kernel evaluation and LUT generation.
// Behavioural restatement of the reference's src/filters.c (pl_filter_sample :82-124,
// filter_cutoffs :126-155, compute_row :159-182, pl_filter_generate :186-245, kernel
// functions :250-560, presets :620-941); written from the formulas, not from the code.
// Parity with the reference is pinned by tests/test_t0_filters.py (bit-exact LUTs).
#define _GNU_SOURCE
#include "../internal.h"

#include <math.h>

// ---- kernel functions: all are even, defined on [0, radius] --------------------------------

#define KFN(name) static double name(const struct pl_filter_ctx *f, double x)

KFN(k_box)      { (void) f; (void) x; return 1.0; }
KFN(k_triangle) { return 1.0 - x / f->radius; }
KFN(k_cosine)   { (void) f; return cos(x); }
KFN(k_hann)     { (void) f; return 0.5 + 0.5 * cos(M_PI * x); }
KFN(k_hamming)  { (void) f; return 0.54 + 0.46 * cos(M_PI * x); }
KFN(k_welch)    { (void) f; return 1.0 - x * x; }

// modified Bessel function of the first kind, order 0, by its power series
static double bessel_i0(double x)
{
    const double q = x * x / 4.0;
    double sum = 1.0, term = q;
    for (int n = 2; term > 1e-12; n++) {
        sum += term;
        term *= q / (n * n);
    }
    return sum;
}

KFN(k_kaiser)
{
    const double alpha = fmax(f->params[0], 0.0);
    return bessel_i0(alpha * sqrt(1.0 - x * x)) / bessel_i0(alpha);
}

KFN(k_blackman)
{
    const double a = f->params[0];
    const double px = M_PI * x;
    return (1 - a) / 2.0 + 1 / 2.0 * cos(px) + a / 2.0 * cos(2 * px);
}

KFN(k_bohman)
{
    (void) f;
    const double px = M_PI * x;
    return (1.0 - x) * cos(px) + sin(px) / M_PI;
}

KFN(k_gaussian) { return exp(-2.0 * x * x / f->params[0]); }

KFN(k_quadratic)
{
    (void) f;
    return x < 0.5 ? 1.0 - 4.0/3.0 * (x * x) : 2.0 / 3.0 * (x - 1.5) * (x - 1.5);
}

KFN(k_sinc)
{
    (void) f;
    if (x < 1e-8)
        return 1.0;
    const double px = M_PI * x;
    return sin(px) / px;
}

KFN(k_jinc)
{
    (void) f;
    if (x < 1e-8)
        return 1.0;
    const double px = M_PI * x;
    return 2.0 * j1(px) / px;
}

KFN(k_sphinx)
{
    (void) f;
    if (x < 1e-8)
        return 1.0;
    const double px = M_PI * x;
    return 3.0 * (sin(px) - px * cos(px)) / (px * px * px);
}

// Mitchell-Netravali family, B = params[0], C = params[1]
KFN(k_cubic)
{
    const double b = f->params[0], c = f->params[1];
    const double p0 = 6.0 - 2.0 * b;
    if (x < 1.0) {
        const double p2 = -18.0 + 12.0 * b + 6.0 * c, p3 = 12.0 - 9.0 * b - 6.0 * c;
        return (p0 + x * x * (p2 + x * p3)) / p0;
    }
    const double q0 = 8.0 * b + 24.0 * c, q1 = -12.0 * b - 48.0 * c,
                 q2 = 6.0 * b + 30.0 * c, q3 = -b - 6.0 * c;
    return (q0 + x * (q1 + x * (q2 + x * q3))) / p0;
}

KFN(k_spline16)
{
    (void) f;
    if (x < 1.0)
        return ((x - 9.0/5.0) * x - 1.0/5.0) * x + 1.0;
    const double t = x - 1;
    return ((-1.0/3.0 * t + 4.0/5.0) * t - 7.0/15.0) * t;
}

KFN(k_spline36)
{
    (void) f;
    if (x < 1.0)
        return ((13.0/11.0 * x - 453.0/209.0) * x - 3.0/209.0) * x + 1.0;
    if (x < 2.0) {
        const double t = x - 1;
        return ((-6.0/11.0 * t + 270.0/209.0) * t - 156.0/209.0) * t;
    }
    const double t = x - 2;
    return ((1.0/11.0 * t - 45.0/209.0) * t + 26.0/209.0) * t;
}

KFN(k_spline64)
{
    (void) f;
    if (x < 1.0)
        return ((49.0/41.0 * x - 6387.0/2911.0) * x - 3.0/2911.0) * x + 1.0;
    if (x < 2.0) {
        const double t = x - 1;
        return ((-24.0/41.0 * t + 4032.0/2911.0) * t - 2328.0/2911.0) * t;
    }
    if (x < 3.0) {
        const double t = x - 2;
        return ((6.0/41.0 * t - 1008.0/2911.0) * t + 582.0/2911.0) * t;
    }
    const double t = x - 3;
    return ((-1.0/41.0 * t + 168.0/2911.0) * t - 97.0/2911.0) * t;
}

#define FN(sym, nm, fn, rad, ...) \
    const struct pl_filter_function sym = { .name = nm, .weight = fn, .radius = rad, __VA_ARGS__ }

FN(pl_filter_function_box,       "box",       k_box,       1.0, .resizable = true);
FN(pl_filter_function_triangle,  "triangle",  k_triangle,  1.0, .resizable = true);
FN(pl_filter_function_cosine,    "cosine",    k_cosine,    M_PI / 2.0);
FN(pl_filter_function_hann,      "hann",      k_hann,      1.0);
FN(pl_filter_function_hamming,   "hamming",   k_hamming,   1.0);
FN(pl_filter_function_welch,     "welch",     k_welch,     1.0);
FN(pl_filter_function_kaiser,    "kaiser",    k_kaiser,    1.0, .params = {2.0}, .tunable = {true});
FN(pl_filter_function_blackman,  "blackman",  k_blackman,  1.0, .params = {0.16}, .tunable = {true});
FN(pl_filter_function_bohman,    "bohman",    k_bohman,    1.0);
FN(pl_filter_function_gaussian,  "gaussian",  k_gaussian,  2.0, .resizable = true, .params = {1.0}, .tunable = {true});
FN(pl_filter_function_quadratic, "quadratic", k_quadratic, 1.5);
FN(pl_filter_function_sinc,      "sinc",      k_sinc,      1.0, .resizable = true);
FN(pl_filter_function_jinc,      "jinc",      k_jinc,      1.2196698912665045, .resizable = true);
FN(pl_filter_function_sphinx,    "sphinx",    k_sphinx,    1.4302966531242027, .resizable = true);
FN(pl_filter_function_cubic,     "cubic",     k_cubic,     2.0, .params = {1.0, 0.0}, .tunable = {true, true});
FN(pl_filter_function_hermite,   "hermite",   k_cubic,     1.0, .params = {0.0, 0.0});
FN(pl_filter_function_spline16,  "spline16",  k_spline16,  2.0);
FN(pl_filter_function_spline36,  "spline36",  k_spline36,  3.0);
FN(pl_filter_function_spline64,  "spline64",  k_spline64,  4.0);

// "oversample" has no kernel: the name selects pl_shader_sample_oversample, params[0] is its
// snapping threshold (filters.c:604-615)
static double k_oversample(const struct pl_filter_ctx *f, double x) { (void) f; (void) x; return 0.0; }
FN(pl_filter_function_oversample, "oversample", k_oversample, 0.0, .params = {0.0}, .tunable = {true}, .opaque = true);

const struct pl_filter_function * const pl_filter_functions[] = {
    &pl_filter_function_box, &pl_filter_function_triangle, &pl_filter_function_cosine,
    &pl_filter_function_hann, &pl_filter_function_hamming, &pl_filter_function_welch,
    &pl_filter_function_kaiser, &pl_filter_function_blackman, &pl_filter_function_bohman,
    &pl_filter_function_gaussian, &pl_filter_function_quadratic, &pl_filter_function_sinc,
    &pl_filter_function_jinc, &pl_filter_function_sphinx, &pl_filter_function_cubic,
    &pl_filter_function_hermite, &pl_filter_function_spline16, &pl_filter_function_spline36,
    &pl_filter_function_spline64, &pl_filter_function_oversample,
    NULL,
};
const int pl_num_filter_functions = PLB_ARRAY_SIZE(pl_filter_functions) - 1;

const struct pl_filter_function *pl_find_filter_function(const char *name)
{
    if (!name)
        return NULL;
    for (int i = 0; i < pl_num_filter_functions; i++) {
        if (strcmp(name, pl_filter_functions[i]->name) == 0)
            return pl_filter_functions[i];
    }
    return NULL;
}

// ---- presets --------------------------------------------------------------------------------

#define JINC_ZERO3 3.2383154841662362076499
#define JINC_ZERO4 4.2410628637960698819573
#define ROBIDOUX_B      (12 / (19 + 9 * M_SQRT2))
#define ROBIDOUX_C      (113 / (58 + 216 * M_SQRT2))
#define ROBIDOUXSHARP_B (6 / (13 + 7 * M_SQRT2))
#define ROBIDOUXSHARP_C (7 / (2 + 12 * M_SQRT2))
#define UP   PL_FILTER_UPSCALING
#define DOWN PL_FILTER_DOWNSCALING
#define MIX  PL_FILTER_FRAME_MIXING
#define SCALE PL_FILTER_SCALING
#define ALL  PL_FILTER_ALL

#define CFG(sym, nm, desc, kern, allow, rec, ...) \
    const struct pl_filter_config sym = { .name = nm, .description = desc, \
        .kernel = &pl_filter_function_##kern, .allowed = allow, .recommended = rec, __VA_ARGS__ }

CFG(pl_filter_spline16, "spline16", "Spline (2 taps)", spline16, ALL, 0);
CFG(pl_filter_spline36, "spline36", "Spline (3 taps)", spline36, ALL, 0);
CFG(pl_filter_spline64, "spline64", "Spline (4 taps)", spline64, ALL, 0);
CFG(pl_filter_nearest, "nearest", "Nearest neighbor", box, UP, UP, .radius = 0.5);
CFG(pl_filter_box, "box", "Box averaging", box, SCALE, DOWN, .radius = 0.5);
CFG(pl_filter_bilinear, "bilinear", "Bilinear", triangle, ALL, SCALE);
CFG(pl_filter_gaussian, "gaussian", "Gaussian", gaussian, ALL, SCALE, .params = {1.0});
CFG(pl_filter_oversample, "oversample", "Oversampling", oversample, UP | MIX, UP | MIX, .params = {0.0});
CFG(pl_filter_sinc, "sinc", "Sinc (unwindowed)", sinc, ALL, 0, .radius = 2.0);
CFG(pl_filter_lanczos, "lanczos", "Lanczos", sinc, ALL, SCALE,
    .window = &pl_filter_function_sinc, .radius = 3.0);
CFG(pl_filter_ginseng, "ginseng", "Ginseng (Jinc-Sinc)", sinc, ALL, 0,
    .window = &pl_filter_function_jinc, .radius = 3.0);
CFG(pl_filter_ewa_jinc, "ewa_jinc", "EWA Jinc (unwindowed)", jinc, SCALE, 0,
    .radius = JINC_ZERO3, .polar = true);
CFG(pl_filter_ewa_lanczos, "ewa_lanczos", "Jinc (EWA Lanczos)", jinc, SCALE, UP,
    .window = &pl_filter_function_jinc, .radius = JINC_ZERO3, .polar = true);
CFG(pl_filter_ewa_lanczossharp, "ewa_lanczossharp", "Sharpened Jinc", jinc, SCALE, UP,
    .window = &pl_filter_function_jinc, .radius = JINC_ZERO3,
    .blur = 0.98125058372237073562493, .polar = true);
CFG(pl_filter_ewa_lanczos4sharpest, "ewa_lanczos4sharpest", "Sharpened Jinc-AR, 4 taps", jinc,
    SCALE, UP, .window = &pl_filter_function_jinc, .radius = JINC_ZERO4,
    .blur = 0.88451209326050047745788, .antiring = 0.8, .polar = true);
CFG(pl_filter_ewa_ginseng, "ewa_ginseng", "EWA Ginseng", jinc, SCALE, 0,
    .window = &pl_filter_function_sinc, .radius = JINC_ZERO3, .polar = true);
CFG(pl_filter_ewa_hann, "ewa_hann", "EWA Hann", jinc, SCALE, 0,
    .window = &pl_filter_function_hann, .radius = JINC_ZERO3, .polar = true);
CFG(pl_filter_bicubic, "bicubic", "Bicubic", cubic, SCALE, SCALE, .params = {1.0, 0.0});
CFG(pl_filter_hermite, "hermite", "Hermite", hermite, ALL, DOWN | MIX);
CFG(pl_filter_catmull_rom, "catmull_rom", "Catmull-Rom", cubic, ALL, SCALE, .params = {0.0, 0.5});
CFG(pl_filter_mitchell, "mitchell", "Mitchell-Netravali", cubic, ALL, DOWN,
    .params = {1/3.0, 1/3.0});
CFG(pl_filter_mitchell_clamp, "mitchell_clamp", "Mitchell (clamped)", cubic, ALL, 0,
    .params = {1/3.0, 1/3.0}, .clamp = 1.0);
CFG(pl_filter_robidoux, "robidoux", "Robidoux", cubic, ALL, 0,
    .params = {ROBIDOUX_B, ROBIDOUX_C});
CFG(pl_filter_robidouxsharp, "robidouxsharp", "RobidouxSharp", cubic, ALL, 0,
    .params = {ROBIDOUXSHARP_B, ROBIDOUXSHARP_C});
CFG(pl_filter_ewa_robidoux, "ewa_robidoux", "EWA Robidoux", cubic, SCALE, 0,
    .params = {ROBIDOUX_B, ROBIDOUX_C}, .polar = true);
CFG(pl_filter_ewa_robidouxsharp, "ewa_robidouxsharp", "EWA RobidouxSharp", cubic, SCALE, 0,
    .params = {ROBIDOUXSHARP_B, ROBIDOUXSHARP_C}, .polar = true);

const struct pl_filter_config * const pl_filter_configs[] = {
    &pl_filter_bilinear, &pl_filter_nearest, &pl_filter_spline16, &pl_filter_spline36,
    &pl_filter_spline64, &pl_filter_lanczos, &pl_filter_ewa_lanczos,
    &pl_filter_ewa_lanczossharp, &pl_filter_ewa_lanczos4sharpest, &pl_filter_bicubic,
    &pl_filter_hermite, &pl_filter_gaussian, &pl_filter_oversample, &pl_filter_mitchell, &pl_filter_mitchell_clamp,
    &pl_filter_sinc, &pl_filter_ginseng, &pl_filter_ewa_jinc, &pl_filter_ewa_ginseng,
    &pl_filter_ewa_hann, &pl_filter_catmull_rom, &pl_filter_robidoux,
    &pl_filter_robidouxsharp, &pl_filter_ewa_robidoux, &pl_filter_ewa_robidouxsharp,
    NULL,
};
const int pl_num_filter_configs = PLB_ARRAY_SIZE(pl_filter_configs) - 1;

const struct pl_filter_config *pl_find_filter_config(const char *name, enum pl_filter_usage usage)
{
    if (!name)
        return NULL;
    for (int i = 0; i < pl_num_filter_configs; i++) {
        const struct pl_filter_config *c = pl_filter_configs[i];
        if ((c->allowed & usage) == usage && strcmp(name, c->name) == 0)
            return c;
    }
    return NULL;
}

// ---- evaluation -------------------------------------------------------------------------------

static float radius_bound(const struct pl_filter_config *c)
{
    const float r = c->radius && c->kernel->resizable ? c->radius : c->kernel->radius;
    return c->blur > 0.0 ? r * c->blur : r;
}

// src/filters.h pl_filter_radius_bound (library-internal in the reference; the frame mixer uses it)
float plb_filter_radius_bound(const struct pl_filter_config *c)
{
    return radius_bound(c);
}

static bool same_function(const struct pl_filter_function *a, const struct pl_filter_function *b)
{
    return (a ? a->weight : NULL) == (b ? b->weight : NULL);
}

bool pl_filter_config_eq(const struct pl_filter_config *a, const struct pl_filter_config *b)
{
    if (!a || !b)
        return a == b;
    if (!same_function(a->kernel, b->kernel) || !same_function(a->window, b->window) ||
        a->radius != b->radius || a->clamp != b->clamp || a->blur != b->blur ||
        a->taper != b->taper || a->polar != b->polar || a->antiring != b->antiring)
        return false;
    for (int i = 0; i < PL_FILTER_MAX_PARAMS; i++) {
        if (a->kernel->tunable[i] != b->kernel->tunable[i])
            return false;
        if (a->kernel->tunable[i] && a->params[i] != b->params[i])
            return false;
        if (!a->window)
            continue;
        if (a->window->tunable[i] != b->window->tunable[i])
            return false;
        if (a->window->tunable[i] && a->wparams[i] != b->wparams[i])
            return false;
    }
    return true;
}

static struct pl_filter_ctx make_ctx(const struct pl_filter_function *fn, float radius,
                                     const float *user)
{
    struct pl_filter_ctx ctx = { .radius = radius };
    for (int i = 0; i < PL_FILTER_MAX_PARAMS; i++)
        ctx.params[i] = fn->tunable[i] ? user[i] : fn->params[i];
    return ctx;
}

double pl_filter_sample(const struct pl_filter_config *c, double x)
{
    const float radius = radius_bound(c);
    x = fabs(x);
    if (x > radius)
        return 0.0;

    // taper flattens the top of the kernel, blur stretches it
    double kx = x <= c->taper ? 0.0 : (x - c->taper) / (1.0 - c->taper / radius);
    if (c->blur > 0.0)
        kx /= c->blur;

    const struct pl_filter_ctx kctx = make_ctx(c->kernel, radius, c->params);
    double k = c->kernel->weight(&kctx, kx);
    if (c->window) {
        // the window is stretched to cover exactly [0, radius]
        const struct pl_filter_ctx wctx = make_ctx(c->window, c->window->radius, c->wparams);
        k *= c->window->weight(&wctx, x / radius * c->window->radius);
    }
    return k < 0 ? (1 - c->clamp) * k : k;
}

// Scan outwards in steps of 0.01 for the places where |f| falls through `cutoff`; the first
// is the main lobe, the last the effective radius (refined by one secant step each).
static void find_cutoffs(const struct pl_filter_config *c, float cutoff, float *radius,
                         float *radius_zero)
{
    const float bound = radius_bound(c);
    const float step = 1e-2f;
    float prev = 0.0, fprev = pl_filter_sample(c, prev);
    bool found = false;
    for (float x = 0.0; x < bound + step; x += step) {
        const float fx = pl_filter_sample(c, x);
        const bool down = fprev > cutoff && fx <= cutoff;
        const bool up = fprev < -cutoff && fx >= -cutoff;
        if (down || up) {
            float root = x - fx * (x - prev) / (fx - fprev);
            root = fminf(root, bound);
            *radius = root;
            if (!found)
                *radius_zero = root;
            found = true;
        }
        prev = x;
        fprev = fx;
    }
    if (!found)
        *radius = *radius_zero = bound;
}

struct filter_priv {
    struct pl_filter_t pub;
    struct pl_filter_function kernel, window;
    float *weights;
};

pl_filter pl_filter_generate(pl_log log, const struct pl_filter_params *params)
{
    if (!params || params->lut_entries <= 0 || !params->config.kernel) {
        plb_msg(log, PL_LOG_FATAL, "Invalid params: missing lut_entries or config.kernel");
        return NULL;
    }
    if (params->config.kernel->opaque || (params->config.window && params->config.window->opaque)) {
        PLB_ERR(log, "Trying to use an opaque filter function in non-opaque context!");
        return NULL;
    }

    struct filter_priv *p = calloc(1, sizeof(*p));
    struct pl_filter_t *f = &p->pub;
    f->params = *params;
    p->kernel = *params->config.kernel;
    f->params.config.kernel = &p->kernel;
    if (params->config.window) {
        p->window = *params->config.window;
        f->params.config.window = &p->window;
    }
    find_cutoffs(&params->config, params->cutoff, &f->radius, &f->radius_zero);

    const int n = params->lut_entries;
    if (params->config.polar) {
        // 1-D table indexed by distance in [0, radius]
        p->weights = malloc(n * sizeof(float));
        for (int i = 0; i < n; i++)
            p->weights[i] = pl_filter_sample(&params->config, f->radius * i / (n - 1));
    } else {
        f->row_size = ceilf(f->radius) * 2;
        if (params->max_row_size && f->row_size > params->max_row_size) {
            PLB_INFO(log, "Required filter size %d exceeds the maximum allowed size of %d.",
                     f->row_size, params->max_row_size);
            f->row_size = params->max_row_size;
            f->insufficient = true;
        }
        const int align = PLB_MAX(params->row_stride_align, 1);
        f->row_stride = (f->row_size + align - 1) / align * align;
        p->weights = calloc((size_t) n * f->row_stride, sizeof(float));
        // one normalised row of taps per sub-pixel phase
        for (int i = 0; i < n; i++) {
            float *row = p->weights + (size_t) f->row_stride * i;
            const double center = f->row_size / 2 - 1 + i / (double) (n - 1);
            double wsum = 0.0;
            for (int k = 0; k < f->row_size; k++) {
                const double w = pl_filter_sample(&params->config, k - center);
                row[k] = w;
                wsum += w;
            }
            for (int k = 0; k < f->row_size; k++)
                row[k] /= wsum;
        }
    }
    f->weights = p->weights;
    return f;
}

void pl_filter_free(pl_filter *filter)
{
    if (!filter || !*filter)
        return;
    struct filter_priv *p = (struct filter_priv *) *filter;
    free(p->weights);
    free(p);
    *filter = NULL;
}

// Name -> configuration list of the older scaler-selection API (filters.h:316-329; what
// vf_libplacebo and mpv's option tables walk): recommended filters first, then the rest, then
// aliases without a description. "none" = the built-in texture sampling.
const struct pl_filter_preset pl_filter_presets[] = {
    { "none",                 NULL,                              "Built-in sampling" },
    { "bilinear",             &pl_filter_bilinear,               "Bilinear" },
    { "nearest",              &pl_filter_nearest,                "Nearest neighbour" },
    { "bicubic",              &pl_filter_bicubic,                "Bicubic" },
    { "lanczos",              &pl_filter_lanczos,                "Lanczos" },
    { "ewa_lanczos",          &pl_filter_ewa_lanczos,            "Jinc (EWA Lanczos)" },
    { "ewa_lanczossharp",     &pl_filter_ewa_lanczossharp,       "Sharpened Jinc" },
    { "ewa_lanczos4sharpest", &pl_filter_ewa_lanczos4sharpest,   "Sharpened Jinc-AR, 4 taps" },
    { "gaussian",             &pl_filter_gaussian,               "Gaussian" },
    { "spline16",             &pl_filter_spline16,               "Spline (2 taps)" },
    { "spline36",             &pl_filter_spline36,               "Spline (3 taps)" },
    { "spline64",             &pl_filter_spline64,               "Spline (4 taps)" },
    { "mitchell",             &pl_filter_mitchell,               "Mitchell-Netravali" },
    { "sinc",                 &pl_filter_sinc,                   "Sinc (unwindowed)" },
    { "ginseng",              &pl_filter_ginseng,                "Ginseng (Jinc-Sinc)" },
    { "ewa_jinc",             &pl_filter_ewa_jinc,               "EWA Jinc (unwindowed)" },
    { "ewa_ginseng",          &pl_filter_ewa_ginseng,            "EWA Ginseng" },
    { "ewa_hann",             &pl_filter_ewa_hann,               "EWA Hann" },
    { "hermite",              &pl_filter_hermite,                "Hermite" },
    { "catmull_rom",          &pl_filter_catmull_rom,            "Catmull-Rom" },
    { "robidoux",             &pl_filter_robidoux,               "Robidoux" },
    { "robidouxsharp",        &pl_filter_robidouxsharp,          "RobidouxSharp" },
    { "ewa_robidoux",         &pl_filter_ewa_robidoux,           "EWA Robidoux" },
    { "ewa_robidouxsharp",    &pl_filter_ewa_robidouxsharp,      "EWA RobidouxSharp" },
    { "triangle",             &pl_filter_bilinear,               NULL },
    { "ewa_hanning",          &pl_filter_ewa_hann,               NULL },
    { 0 },
};
const int pl_num_filter_presets = PLB_ARRAY_SIZE(pl_filter_presets) - 1;

const struct pl_filter_preset *pl_find_filter_preset(const char *name)
{
    for (int i = 0; name && i < pl_num_filter_presets; i++) {
        if (strcmp(pl_filter_presets[i].name, name) == 0)
            return &pl_filter_presets[i];
    }
    return NULL;
}
