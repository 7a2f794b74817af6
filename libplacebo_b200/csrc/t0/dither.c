// Host-side dither matrices and error-diffusion kernel tables.
// Behavioural restatement of the reference's src/dither.c: recursive Bayer matrix (:36-55),
// void-and-cluster blue noise with 64-bit fixed-point energy on a torus (:57-190), classic
// error diffusion kernels (:192-290). Pinned bit-exactly by tests/test_t0_dither.py.
//
// The reference breaks ties with libc rand() (dither.c:162), i.e. its matrix depends on how
// often the process called rand() before. We reproduce what a fresh process produces by
// carrying a private copy of glibc's default TYPE_3 additive-feedback generator seeded with 1,
// so the matrix is deterministic and equal to the reference's first-call result.
#include "../internal.h"

#include <math.h>

void pl_generate_bayer_matrix(float *data, int size)
{
    // M(2n) = [[4M, 4M+2], [4M+3, 4M+1]] / 4, built in place quadrant by quadrant
    data[0] = 0;
    for (int n = 1; n < size; n *= 2) {
        const double step = 4.0 * n * n;
        for (int y = 0; y < n; y++) {
            for (int x = 0; x < n; x++) {
                const float base = data[y * size + x];
                data[(y + n) * size + (x + n)] = base + 1 / step;
                data[y * size + (x + n)] = base + 2 / step;
                data[(y + n) * size + x] = base + 3 / step;
            }
        }
    }
}

// glibc random_r() TYPE_3 (x^31 + x^3 + 1), as seeded by srand(1)
struct glibc_rand {
    int32_t r[34 + 310];
    int n;
};

static void glibc_srand1(struct glibc_rand *g)
{
    int32_t *r = g->r;
    r[0] = 1;
    for (int i = 1; i < 31; i++) {
        // 16807 * r[i-1] mod (2^31 - 1) without overflow (Schrage)
        const int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t word = 16807 * lo - 2836 * hi;
        if (word < 0)
            word += 2147483647;
        r[i] = (int32_t) word;
    }
    for (int i = 31; i < 34; i++)
        r[i] = r[i - 31];
    for (int i = 34; i < 344; i++)
        r[i] = (int32_t) ((uint32_t) r[i - 31] + (uint32_t) r[i - 3]);
    g->n = 0;
}

static int glibc_rand_next(struct glibc_rand *g)
{
    // sliding window over the last 34 values
    int32_t *r = g->r;
    const int i = 344 + g->n;
    const uint32_t v = (uint32_t) r[(i - 31) % 344] + (uint32_t) r[(i - 3) % 344];
    r[i % 344] = (int32_t) v;
    g->n++;
    return (int) (v >> 1);
}

struct vc_ctx {
    unsigned sizeb, size, size2;
    uint64_t *kernel;   // gaussian energy stamp, size2 entries, centred at `middle`
    uint64_t *energy;   // accumulated energy per cell
    bool *placed;
    uint32_t *rank;     // order in which cells were placed
    uint32_t *ties;
    unsigned middle;
};

#define CELL(k, x, y) ((x) | ((y) << (k)->sizeb))

void pl_generate_blue_noise(float *data, int size)
{
    unsigned sizeb = 0;
    while ((1 << sizeb) < size)
        sizeb++;
    if (size <= 0 || (1 << sizeb) != size || sizeb < 1 || sizeb > 8)
        return;

    struct vc_ctx k = { .sizeb = sizeb, .size = size, .size2 = (unsigned) size * size };
    k.kernel = calloc(k.size2, sizeof(uint64_t));
    k.energy = calloc(k.size2, sizeof(uint64_t));
    k.placed = calloc(k.size2, sizeof(bool));
    k.rank = calloc(k.size2, sizeof(uint32_t));
    k.ties = calloc(k.size2, sizeof(uint32_t));

    // exponential-falloff stamp of diameter size-1, scaled so that the sum of a full matrix
    // of stamps still fits 64 bits
    const unsigned radius = k.size / 2 - 1;
    const unsigned diam = radius * 2 + 1, diam2 = diam * diam;
    k.middle = CELL(&k, radius, radius);
    const double sigma = -log(1.5 / (double) UINT64_MAX * diam2) / radius;
    for (unsigned gy = 0; gy <= radius; gy++) {
        for (unsigned gx = 0; gx <= gy; gx++) {
            const int cx = (int) gx - radius, cy = (int) gy - radius;
            const int sq = cx * cx + cy * cy;
            const double e = exp(-sqrt(sq) * sigma);
            const uint64_t v = e / diam2 * (double) UINT64_MAX;
            // 8-fold symmetry
            const unsigned mx = diam - 1 - gx, my = diam - 1 - gy;
            k.kernel[CELL(&k, gx, gy)] = k.kernel[CELL(&k, gy, gx)] =
            k.kernel[CELL(&k, gx, my)] = k.kernel[CELL(&k, gy, mx)] =
            k.kernel[CELL(&k, mx, gy)] = k.kernel[CELL(&k, my, gx)] =
            k.kernel[CELL(&k, mx, my)] = k.kernel[CELL(&k, my, mx)] = v;
        }
    }

    struct glibc_rand rng;
    glibc_srand1(&rng);

    for (unsigned n = 0; n < k.size2; n++) {
        // the emptiest (lowest-energy) free cell; ties resolved uniformly at random
        uint64_t best = UINT64_MAX;
        unsigned nties = 0;
        for (unsigned c = 0; c < k.size2; c++) {
            if (k.placed[c])
                continue;
            const uint64_t e = k.energy[c];
            if (e <= best) {
                if (e != best) {
                    best = e;
                    nties = 0;
                }
                k.ties[nties++] = c;
            }
        }
        unsigned pick;
        if (nties == 1)
            pick = k.ties[0];
        else if (nties == k.size2)
            pick = k.size2 / 2;
        else
            pick = k.ties[glibc_rand_next(&rng) % nties];

        // stamp the kernel centred on `pick`, wrapping in linear index space
        k.placed[pick] = true;
        const unsigned start = (k.middle + k.size2 - pick) & (k.size2 - 1);
        for (unsigned m = 0; m < k.size2; m++)
            k.energy[m] += k.kernel[(start + m) & (k.size2 - 1)];
        k.rank[pick] = n;
    }

    const float invscale = k.size2;
    for (unsigned y = 0; y < k.size; y++)
        for (unsigned x = 0; x < k.size; x++)
            data[x + y * k.size] = k.rank[CELL(&k, x, y)] / invscale;

    free(k.kernel);
    free(k.energy);
    free(k.placed);
    free(k.rank);
    free(k.ties);
}

#define EDK(sym, nm, desc, sh, div, ...) \
    const struct pl_error_diffusion_kernel sym = { .name = nm, .description = desc, \
        .shift = sh, .divisor = div, .pattern = __VA_ARGS__ }

// pattern rows: dy = 0..2, columns dx = -2..+2; (0,0) is the pixel being quantised
EDK(pl_error_diffusion_simple, "simple", "Simple error diffusion", 1, 2,
    {{0, 0, 0, 1, 0}, {0, 0, 1, 0, 0}, {0, 0, 0, 0, 0}});
EDK(pl_error_diffusion_false_fs, "false-fs", "False Floyd-Steinberg kernel", 1, 8,
    {{0, 0, 0, 3, 0}, {0, 0, 3, 2, 0}, {0, 0, 0, 0, 0}});
EDK(pl_error_diffusion_sierra_lite, "sierra-lite", "Sierra Lite kernel", 2, 4,
    {{0, 0, 0, 2, 0}, {0, 1, 1, 0, 0}, {0, 0, 0, 0, 0}});
EDK(pl_error_diffusion_floyd_steinberg, "floyd-steinberg", "Floyd Steinberg kernel", 2, 16,
    {{0, 0, 0, 7, 0}, {0, 3, 5, 1, 0}, {0, 0, 0, 0, 0}});
EDK(pl_error_diffusion_atkinson, "atkinson", "Atkinson kernel", 2, 8,
    {{0, 0, 0, 1, 1}, {0, 1, 1, 1, 0}, {0, 0, 1, 0, 0}});
EDK(pl_error_diffusion_jarvis_judice_ninke, "jarvis-judice-ninke",
    "Jarvis, Judice & Ninke kernel", 3, 48,
    {{0, 0, 0, 7, 5}, {3, 5, 7, 5, 3}, {1, 3, 5, 3, 1}});
EDK(pl_error_diffusion_stucki, "stucki", "Stucki kernel", 3, 42,
    {{0, 0, 0, 8, 4}, {2, 4, 8, 4, 2}, {1, 2, 4, 2, 1}});
EDK(pl_error_diffusion_burkes, "burkes", "Burkes kernel", 3, 32,
    {{0, 0, 0, 8, 4}, {2, 4, 8, 4, 2}, {0, 0, 0, 0, 0}});
EDK(pl_error_diffusion_sierra2, "sierra-2", "Two-row Sierra", 3, 16,
    {{0, 0, 0, 4, 3}, {1, 2, 3, 2, 1}, {0, 0, 0, 0, 0}});
EDK(pl_error_diffusion_sierra3, "sierra-3", "Three-row Sierra", 3, 32,
    {{0, 0, 0, 5, 3}, {2, 4, 5, 4, 2}, {0, 2, 3, 2, 0}});

const struct pl_error_diffusion_kernel * const pl_error_diffusion_kernels[] = {
    &pl_error_diffusion_simple, &pl_error_diffusion_false_fs, &pl_error_diffusion_sierra_lite,
    &pl_error_diffusion_floyd_steinberg, &pl_error_diffusion_atkinson,
    &pl_error_diffusion_jarvis_judice_ninke, &pl_error_diffusion_stucki,
    &pl_error_diffusion_burkes, &pl_error_diffusion_sierra2, &pl_error_diffusion_sierra3,
    NULL
};
const int pl_num_error_diffusion_kernels = PLB_ARRAY_SIZE(pl_error_diffusion_kernels) - 1;

const struct pl_error_diffusion_kernel *pl_find_error_diffusion_kernel(const char *name)
{
    for (int i = 0; name && i < pl_num_error_diffusion_kernels; i++) {
        if (strcmp(name, pl_error_diffusion_kernels[i]->name) == 0)
            return pl_error_diffusion_kernels[i];
    }
    return NULL;
}
