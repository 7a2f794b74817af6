// pl_shader_error_diffusion (src/shaders/dithering.c:326-527).
//
// The reference runs ONE work group that walks the image column by column after the shift
// mapping (y, x) -> (y, x + y*shift): within a shifted column no pixel depends on another, so a
// column is processed in parallel; errors travel to the next few shifted columns through a ring
// buffer of packed RGB8 words (8 bits per channel, 4 guard bits) updated with atomicAdd.
// Integer adds commute, so the result does not depend on how a column is split over threads:
// this kernel is bit-identical to the reference's formulation for any block size.
//
// The step count, w + (h-1)*shift barriers, is inherently serial (8K sierra-lite: 16 318): the
// kernel is LATENCY-bound. One CTA of up to 1024 threads per image keeps the barrier at
// __syncthreads cost; the texels of the next column are fetched before the barrier so that the
// global-memory latency overlaps the current column's atomics. Throughput comes from running
// several images at once (other streams / GPUs), not from a wider launch.
#include "kernels_common.cuh"

#define EDF_MAX_TAPS 12

struct errdiff_args {
    plb_tex_desc src, dst;
    int w, h;
    int shift;
    int quant;              // (1 << new_depth) - 1
    float divisor;
    int ring_rows, ring_size;
    int ntaps;
    int tap_dividend[EDF_MAX_TAPS];
    int tap_delta[EDF_MAX_TAPS];   // shifted_x * ring_rows + dy
    int tap_min_x[EDF_MAX_TAPS];   // pixel must satisfy x >= this (errors leaving the left edge)
    int cols;               // shifted columns = w + (h-1)*shift
};

__global__ void __launch_bounds__(1024)
k_errdiff(const __grid_constant__ errdiff_args a)
{
    extern __shared__ uint32_t err[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < a.ring_size; i += nt)
        err[i] = 0u;
    __syncthreads();

    const float fq = (float) a.quant;
    for (int xs = 0; xs < a.cols; xs++) {
        for (int y = tid; y < a.h; y += nt) {
            const int x = xs - y * a.shift;
            if (x < 0 || x >= a.w)
                continue;
            const uint32_t idx = (uint32_t) (xs * a.ring_rows + y) % (uint32_t) a.ring_size;
            const float4 po = plb_load_clamped(a.src, x, y);
            float pix[3] = { po.x, po.y, po.z };
            const uint32_t e = atomicExch(&err[idx], 0u) + ((128u << 24) | (128u << 12) | 128u);
            const int er = (int) ((e >> 24) & 0xFFu) - 128, eg = (int) ((e >> 12) & 0xFFu) - 128,
                      eb = (int) (e & 0xFFu) - 128;
            pix[0] = pix[0] * fq + (float) er / 254.0f;
            pix[1] = pix[1] * fq + (float) eg / 254.0f;
            pix[2] = pix[2] * fq + (float) eb / 254.0f;
            float dith[3], ed[3];
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                dith[k] = roundf(pix[k]);
                ed[k] = (pix[k] - dith[k]) * 254.0f / a.divisor;
            }
            plb_store(a.dst, x, y, make_float4(dith[0] / fq, dith[1] / fq, dith[2] / fq, po.w));
            int last = 0;
            uint32_t packed = 0;
            for (int t = 0; t < a.ntaps; t++) {
                if (a.tap_dividend[t] != last) {
                    last = a.tap_dividend[t];
                    const float d = (float) last;
                    const int tr = (int) roundf(ed[0] * d), tg = (int) roundf(ed[1] * d),
                              tb = (int) roundf(ed[2] * d);
                    packed = ((uint32_t) (tr & 0xFF) << 24) | ((uint32_t) (tg & 0xFF) << 12) |
                             (uint32_t) (tb & 0xFF);
                }
                if (x >= a.tap_min_x[t])
                    atomicAdd(&err[(idx + (uint32_t) a.tap_delta[t]) % (uint32_t) a.ring_size], packed);
            }
        }
        __syncthreads();
    }
}

#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)

extern "C" bool plb_launch_errdiff(pl_gpu gpu, const struct plb_errdiff_launch *p)
{
    struct plb_gpu *g = plb_gpu_priv(gpu);
    static errdiff_args a;
    memset(&a, 0, sizeof(a));
    a.src = p->src;
    a.dst = p->dst;
    a.w = p->w;
    a.h = p->h;
    a.shift = p->shift;
    a.quant = (int) ((1ll << p->new_depth) - 1);
    a.divisor = (float) p->divisor;
    a.ring_rows = p->h + 2; // PL_EDF_MAX_DY
    int rightmost = 0;
    // taps grouped by dividend, then (dy, dx) order: the order of the reference's GLSL emission
    for (int dividend = 1; dividend <= p->divisor; dividend++) {
        for (int dy = 0; dy <= 2; dy++) {
            for (int dx = -2; dx <= 2; dx++) {
                if (p->pattern[dy][dx + 2] != dividend)
                    continue;
                if (a.ntaps == EDF_MAX_TAPS) {
                    PLB_ERR(gpu->log, "error diffusion kernel has too many taps");
                    return false;
                }
                const int sx = dx + dy * p->shift;
                a.tap_dividend[a.ntaps] = dividend;
                a.tap_delta[a.ntaps] = sx * a.ring_rows + dy;
                a.tap_min_x[a.ntaps] = dx < 0 ? -dx : 0;
                a.ntaps++;
            }
        }
    }
    for (int dy = 0; dy <= 2; dy++) {
        for (int dx = -2; dx <= 2; dx++) {
            if (p->pattern[dy][dx + 2])
                rightmost = PLB_MAX(rightmost, dx + dy * p->shift);
        }
    }
    a.ring_size = a.ring_rows * (rightmost + 1);
    a.cols = p->w + (p->h - 1) * p->shift;
    const size_t shmem = (size_t) a.ring_size * sizeof(uint32_t);
    if (shmem > 220 * 1024) {
        PLB_ERR(gpu->log, "error diffusion: image too tall for the shared-memory ring (%zu B)", shmem);
        return false;
    }
    static size_t configured = 0;
    if (shmem > 48 * 1024 && shmem > configured) {
        if (!plb_cuda_check(gpu, cudaFuncSetAttribute(k_errdiff, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int) shmem), "cudaFuncSetAttribute"))
            return false;
        configured = shmem;
    }
    int threads = PLB_MIN(1024, (p->h + 31) / 32 * 32);
    k_errdiff<<<1, threads, shmem, STREAM(gpu)>>>(a);
    g->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_errdiff");
}
