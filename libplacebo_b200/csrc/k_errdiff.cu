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
// global-memory latency overlaps the current column's atomics.
//
// Tall images (more rows than one CTA has threads: 4K, 8K) run on a THREAD-BLOCK CLUSTER instead
// (k_errdiff_cluster): up to 8 CTAs on 8 SMs, one image row per thread, the error ring split by
// rows over the CTAs' shared memories; errors that cross a band boundary are added to the
// neighbour's ring through distributed shared memory, and the per-column barrier is the
// hardware cluster barrier. 8K sierra-lite: 92 ms -> see profiles/r02_summary.md.
// Beyond that, throughput comes from running several images at once (other streams / GPUs).
#include "kernels_common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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
    int tap_sx[EDF_MAX_TAPS], tap_dy[EDF_MAX_TAPS]; // shifted column offset (>= 0) and row offset
    int tap_min_x[EDF_MAX_TAPS];   // pixel must satisfy x >= this (errors leaving the left edge)
    int cols;               // shifted columns = w + (h-1)*shift
};

// texel fetch / store of the usual FBO layout (rgba16hf) without the run-time format dispatch
template <bool HALF4>
PLB_DEV float4 edf_load(const plb_tex_desc &t, int x, int y)
{
    if (HALF4) {
        const uint2 raw = *(const uint2 *) ((const char *) t.ptr + (size_t) y * t.pitch + (size_t) x * 8);
        const float2 lo = __half22float2(*(const __half2 *) &raw.x), hi = __half22float2(*(const __half2 *) &raw.y);
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    return plb_load_clamped(t, x, y);
}

template <bool HALF4>
PLB_DEV void edf_store(const plb_tex_desc &t, int x, int y, float4 c)
{
    if (HALF4) {
        const __half2 lo = __floats2half2_rn(c.x, c.y), hi = __floats2half2_rn(c.z, c.w);
        uint2 raw;
        raw.x = *(const uint32_t *) &lo;
        raw.y = *(const uint32_t *) &hi;
        *(uint2 *) ((char *) t.ptr + (size_t) y * t.pitch + (size_t) x * 8) = raw;
    } else {
        plb_store(t, x, y, c);
    }
}

// Rows are dealt to threads as y = tid + j * blockDim (j < NJ). The ring slot of shifted column
// xs is xs mod ring_cols, tracked incrementally (no integer division per pixel), and the texels
// of column xs + 1 are fetched before column xs is processed, so that their global-memory
// latency overlaps the arithmetic and the shared-memory atomics of the current column.
#define EDF_NJ 5
template <bool HALF4>
__global__ void __launch_bounds__(1024)
k_errdiff(const __grid_constant__ errdiff_args a)
{
    extern __shared__ uint32_t err[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < a.ring_size; i += nt)
        err[i] = 0u;
    __syncthreads();

    const float fq = (float) a.quant;
    const int ring_cols = a.ring_size / a.ring_rows;
    float4 nxt[EDF_NJ];
    #pragma unroll
    for (int j = 0; j < EDF_NJ; j++) {
        const int y = tid + j * nt, x = 0 - y * a.shift;
        nxt[j] = (y < a.h && x >= 0 && x < a.w) ? edf_load<HALF4>(a.src, x, y) : make_float4(0, 0, 0, 0);
    }
    int slot = 0; // xs mod ring_cols
    for (int xs = 0; xs < a.cols; xs++) {
        float4 cur[EDF_NJ];
        #pragma unroll
        for (int j = 0; j < EDF_NJ; j++) {
            cur[j] = nxt[j];
            const int y = tid + j * nt, x = xs + 1 - y * a.shift;
            if (y < a.h && x >= 0 && x < a.w && xs + 1 < a.cols)
                nxt[j] = edf_load<HALF4>(a.src, x, y);
        }
        #pragma unroll
        for (int j = 0; j < EDF_NJ; j++) {
            const int y = tid + j * nt;
            const int x = xs - y * a.shift;
            if (y >= a.h || x < 0 || x >= a.w)
                continue;
            const uint32_t idx = (uint32_t) (slot * a.ring_rows + y);
            const float4 po = cur[j];
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
            edf_store<HALF4>(a.dst, x, y, make_float4(dith[0] / fq, dith[1] / fq, dith[2] / fq, po.w));
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
                if (x >= a.tap_min_x[t]) {
                    int c = slot + a.tap_sx[t]; // tap's column slot, (xs + sx) mod ring_cols
                    if (c >= ring_cols)
                        c -= ring_cols;
                    atomicAdd(&err[c * a.ring_rows + y + a.tap_dy[t]], packed);
                }
            }
        }
        __syncthreads();
        if (++slot == ring_cols)
            slot = 0;
    }
}

// ---- cluster version ------------------------------------------------------------------------
// CTA `rank` of the cluster owns rows [rank * R, rank * R + R) (the last one two more, for the
// errors that leave the image at the bottom); its ring holds ring_cols columns of R + 2 words.
template <bool HALF4>
__global__ void __launch_bounds__(1024)
k_errdiff_cluster(const __grid_constant__ errdiff_args a, const int R)
{
    extern __shared__ uint32_t err[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int) cluster.block_rank(), nranks = (int) cluster.num_blocks();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ring_cols = a.ring_size / a.ring_rows;
    const int lrows = R + 2;
    for (int i = tid; i < ring_cols * lrows; i += nt)
        err[i] = 0u;
    cluster.sync();

    const float fq = (float) a.quant;
    const int y = rank * R + tid;
    const bool row_ok = tid < R && y < a.h;
    // ring of the CTA that owns row y + dy, and the row's index in it
    uint32_t *tap_ring[3];
    int tap_row[3];
    #pragma unroll
    for (int dy = 0; dy < 3; dy++) {
        const int ty = y + dy, owner = min(ty / R, nranks - 1);
        tap_ring[dy] = owner == rank ? err : cluster.map_shared_rank(err, owner);
        tap_row[dy] = ty - owner * R;
    }
    float4 nxt = make_float4(0, 0, 0, 0);
    {
        const int x = 0 - y * a.shift;
        if (row_ok && x >= 0 && x < a.w)
            nxt = edf_load<HALF4>(a.src, x, y);
    }
    int slot = 0; // xs mod ring_cols
    for (int xs = 0; xs < a.cols; xs++) {
        const float4 po = nxt;
        {
            const int x = xs + 1 - y * a.shift;
            if (row_ok && x >= 0 && x < a.w && xs + 1 < a.cols)
                nxt = edf_load<HALF4>(a.src, x, y);
        }
        const int x = xs - y * a.shift;
        if (row_ok && x >= 0 && x < a.w) {
            float pix[3] = { po.x, po.y, po.z };
            const uint32_t e = atomicExch(&err[slot * lrows + tid], 0u) + ((128u << 24) | (128u << 12) | 128u);
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
            edf_store<HALF4>(a.dst, x, y, make_float4(dith[0] / fq, dith[1] / fq, dith[2] / fq, po.w));
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
                if (x >= a.tap_min_x[t]) {
                    int c = slot + a.tap_sx[t]; // tap's column slot, (xs + sx) mod ring_cols
                    if (c >= ring_cols)
                        c -= ring_cols;
                    const int dy = a.tap_dy[t];
                    uint32_t *ring = dy == 0 ? tap_ring[0] : dy == 1 ? tap_ring[1] : tap_ring[2];
                    const int row = dy == 0 ? tap_row[0] : dy == 1 ? tap_row[1] : tap_row[2];
                    atomicAdd(&ring[c * lrows + row], packed);
                }
            }
        }
        cluster.sync();
        if (++slot == ring_cols)
            slot = 0;
    }
}

#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)

extern "C" bool plb_launch_errdiff(pl_gpu gpu, const struct plb_errdiff_launch *p)
{
    struct plb_gpu *g = plb_gpu_priv(gpu);
    errdiff_args a;
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
                a.tap_sx[a.ntaps] = sx;
                a.tap_dy[a.ntaps] = dy;
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
    const bool half4 = a.src.kind == PLB_FK_HALF && a.src.comps == 4 && !a.src.ptr2 &&
                       a.dst.kind == PLB_FK_HALF && a.dst.comps == 4;
    // Images taller than one CTA's worth of threads: a cluster of up to 8 CTAs, one row per thread
    // (PLB_ERRDIFF_CLUSTER=0 keeps the single-CTA kernel; A/B runs and the bit-identity test)
    static int use_cluster = -1;
    if (use_cluster < 0) {
        const char *e = getenv("PLB_ERRDIFF_CLUSTER");
        use_cluster = e ? atoi(e) : 1;
    }
    if (use_cluster && p->h > 1024) {
        const int nranks = PLB_MIN(8, (p->h + 1023) / 1024 < 2 ? 2 : 8);
        const int R = (p->h + nranks - 1) / nranks;
        const int cthreads = (R + 31) / 32 * 32;
        const int ring_cols = rightmost + 1;
        const size_t csh = (size_t) ring_cols * (R + 2) * sizeof(uint32_t);
        if (cthreads <= 1024 && csh <= 200 * 1024) {
            void (*kern)(errdiff_args, int) = half4 ? k_errdiff_cluster<true> : k_errdiff_cluster<false>;
            if (csh > 48 * 1024 &&
                !plb_cuda_check(gpu, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          (int) csh), "cudaFuncSetAttribute"))
                return false;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nranks);
            cfg.blockDim = dim3(cthreads);
            cfg.dynamicSmemBytes = csh;
            cfg.stream = STREAM(gpu);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = nranks;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, a, R);
            if (err == cudaSuccess) {
                g->launches++;
                return true;
            }
            cudaGetLastError(); // e.g. no 8-SM cluster slot on a partitioned device: single CTA below
        }
    }
    if (shmem > 48 * 1024) { // (per function and device; cheap enough to repeat per launch)
        if (!plb_cuda_check(gpu, cudaFuncSetAttribute(half4 ? k_errdiff<true> : k_errdiff<false>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      (int) shmem), "cudaFuncSetAttribute"))
            return false;
    }
    // every thread owns at most EDF_NJ rows
    int threads = PLB_MIN(1024, (p->h + 31) / 32 * 32);
    if ((p->h + threads - 1) / threads > EDF_NJ) {
        PLB_ERR(gpu->log, "error diffusion: image too tall (%d rows)", p->h);
        return false;
    }
    if (half4)
        k_errdiff<true><<<1, threads, shmem, STREAM(gpu)>>>(a);
    else
        k_errdiff<false><<<1, threads, shmem, STREAM(gpu)>>>(a);
    g->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_errdiff");
}
