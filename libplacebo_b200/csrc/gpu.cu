// pl_cuda: the CUDA sibling of the reference's pl_gpu backends (src/vulkan/gpu.c, src/dummy.c).
// Textures are pitched linear allocations in HBM (kernels address them directly and do their
// own format conversion/filtering in fp32, so no CUarray / texture objects), buffers are
// device allocations with an optional pinned host mirror, timers are cudaEvent pairs, and
// every operation is ordered on one stream so that callers may reuse frames as soon as a call
// returns (reference contract: renderer.h:725-727).
//
// Validation mirrors the wrappers of src/gpu.c (pl_tex_create :383-470, pl_tex_upload/
// download :704-860): bad parameters -> log + false/NULL, never abort.
#include "internal.h"

#include <cuda_runtime.h>
#include <new>

extern "C" {

// ---------------------------------------------------------------------------------------
// log.h
// ---------------------------------------------------------------------------------------

struct plb_log {
    struct pl_log_t pub;
};

pl_log pl_log_create_365(int api_ver, const struct pl_log_params *params)
{
    (void) api_ver;
    struct plb_log *log = (struct plb_log *) calloc(1, sizeof(*log));
    if (params)
        log->pub.params = *params;
    return &log->pub;
}

void pl_log_destroy(pl_log *plog)
{
    if (!plog || !*plog)
        return;
    free((void *) *plog);
    *plog = NULL;
}

struct pl_log_params pl_log_update(pl_log plog, const struct pl_log_params *params)
{
    static const struct pl_log_params none = {};
    if (!plog)
        return none;
    struct pl_log_t *log = (struct pl_log_t *) plog;
    const struct pl_log_params prev = log->params;
    log->params = params ? *params : none;
    return prev;
}

enum pl_log_level pl_log_level_update(pl_log plog, enum pl_log_level level)
{
    if (!plog)
        return PL_LOG_NONE;
    struct pl_log_t *log = (struct pl_log_t *) plog;
    const enum pl_log_level prev = log->params.log_level;
    log->params.log_level = level;
    return prev;
}

void pl_log_simple(void *stream, enum pl_log_level level, const char *msg)
{
    static const char *prefix[] = { "", "fatal", "error", "warn", "info", "debug", "trace" };
    FILE *h = stream ? (FILE *) stream : (level <= PL_LOG_WARN ? stderr : stdout);
    fprintf(h, "%5s: %s\n", prefix[level], msg);
}

void pl_log_color(void *stream, enum pl_log_level level, const char *msg)
{
    // fatal bright red, error red, warning yellow, info green, debug blue, trace bright black
    static const char *sgr[] = { "0", "31;1", "31", "33", "32", "34", "30;1" };
    FILE *h = stream ? (FILE *) stream : (level <= PL_LOG_WARN ? stderr : stdout);
    fprintf(h, "\033[%sm%s\033[0m\n", sgr[level], msg);
    if (level <= PL_LOG_WARN)
        fflush(h);
}

void plb_msg(pl_log log, enum pl_log_level lev, const char *fmt, ...)
{
    if (!log || !log->params.log_cb || lev > log->params.log_level)
        return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    log->params.log_cb(log->params.log_priv, lev, buf);
}

// ---------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------

bool plb_cuda_check(pl_gpu gpu, int err, const char *what)
{
    if (err == cudaSuccess)
        return true;
    struct plb_gpu *p = plb_gpu_priv(gpu);
    PLB_ERR(gpu->log, "CUDA error in %s: %s", what, cudaGetErrorString((cudaError_t) err));
    // sticky errors mean the context is gone (pl_gpu_is_failed, gpu.h:1398)
    if (err != cudaErrorInvalidValue && err != cudaErrorMemoryAllocation &&
        err != cudaErrorInvalidConfiguration)
        p->failed = true;
    cudaGetLastError();
    return false;
}

#define CK(gpu, call) plb_cuda_check(gpu, (int) (call), #call)
#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)

// Every entry point that allocates, copies or launches first makes the gpu's device current and
// restores the caller's on return: two pl_gpu on different devices may live in one process, and
// the host application may switch the current device between calls (e.g. torch.cuda.set_device).
struct dev_guard {
    int prev;
    bool switched;
    explicit dev_guard(pl_gpu gpu) : prev(-1), switched(false)
    {
        const int want = plb_gpu_priv(gpu)->device;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != want)
            switched = cudaSetDevice(want) == cudaSuccess;
    }
    ~dev_guard()
    {
        if (switched)
            cudaSetDevice(prev);
    }
};
#define DEV_GUARD(gpu) dev_guard dev_guard_(gpu)

// C entry points for the host code (dispatch.c wraps every kernel launch)
extern "C" int plb_gpu_enter(pl_gpu gpu)
{
    int prev = -1;
    const int want = plb_gpu_priv(gpu)->device;
    if (cudaGetDevice(&prev) != cudaSuccess || prev == want)
        return -1;
    return cudaSetDevice(want) == cudaSuccess ? prev : -1;
}

extern "C" void plb_gpu_leave(int prev)
{
    if (prev >= 0)
        cudaSetDevice(prev);
}

void *plb_gpu_staging(pl_gpu gpu, size_t size)
{
    DEV_GUARD(gpu);
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (p->staging_size >= size)
        return p->staging;
    if (p->staging) {
        cudaStreamSynchronize(STREAM(gpu));
        cudaFreeHost(p->staging);
        p->staging = NULL;
        p->staging_size = 0;
    }
    if (!CK(gpu, cudaMallocHost(&p->staging, size)))
        return NULL;
    p->staging_size = size;
    return p->staging;
}

void *plb_dev_alloc(pl_gpu gpu, size_t size)
{
    DEV_GUARD(gpu);
    void *ptr = NULL;
    if (!CK(gpu, cudaMalloc(&ptr, size)))
        return NULL;
    return ptr;
}

void plb_dev_free(pl_gpu gpu, void *ptr)
{
    DEV_GUARD(gpu);
    if (!ptr)
        return;
    // work that still reads the allocation is ordered before the free on the same stream
    CK(gpu, cudaFreeAsync(ptr, STREAM(gpu)));
}

} // extern "C"

// Small host <-> device payloads (LUTs, per-frame statistics) move through a copy KERNEL that
// reads / writes the pinned staging area directly (unified addressing) instead of the DMA
// engines: a 1 KB tone-curve upload or the 3264-byte peak statistics would otherwise queue
// behind the frame-sized transfers of the asynchronous upload / download streams on the same
// copy engine, and the render stream (resp. the host) would wait for a whole frame's DMA.
#define PLB_SMALL_XFER (256u << 10)
#define PLB_UP_SLOTS 8
#define PLB_UP_SLOT_BYTES (128u << 10)

__global__ void k_copy_bytes(unsigned char *dst, const unsigned char *src, size_t size)
{
    const size_t nvec = (((uintptr_t) dst | (uintptr_t) src) & 15) ? 0 : size / 16;
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t) gridDim.x * blockDim.x;
    for (size_t i = tid; i < nvec; i += nt)
        ((uint4 *) dst)[i] = ((const uint4 *) src)[i];
    for (size_t i = nvec * 16 + tid; i < size; i += nt)
        dst[i] = src[i];
}

extern "C" {

static bool small_copy(pl_gpu gpu, void *dst, const void *src, size_t size)
{
    const int blocks = (int) PLB_MIN((size / 16 + 255) / 256 + 1, (size_t) 64);
    k_copy_bytes<<<blocks, 256, 0, STREAM(gpu)>>>((unsigned char *) dst, (const unsigned char *) src, size);
    return plb_cuda_check(gpu, cudaGetLastError(), "k_copy_bytes");
}

bool plb_dev_upload(pl_gpu gpu, void *dst, const void *src, size_t size)
{
    DEV_GUARD(gpu);
    // LUT-sized payloads: go through the pinned staging area so that the source may be
    // freed/modified by the caller right after we return
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (size <= PLB_UP_SLOT_BYTES) {
        // small payload: next slot of the pinned ring; wait only for that slot's previous copy
        if (!p->up_ring) {
            if (!CK(gpu, cudaMallocHost(&p->up_ring, (size_t) PLB_UP_SLOTS * PLB_UP_SLOT_BYTES)))
                return false;
            for (int i = 0; i < PLB_UP_SLOTS; i++) {
                cudaEvent_t e;
                if (!CK(gpu, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)))
                    return false;
                p->up_event[i] = e;
            }
        }
        const int slot = p->up_next;
        p->up_next = (slot + 1) % PLB_UP_SLOTS;
        if (!CK(gpu, cudaEventSynchronize((cudaEvent_t) p->up_event[slot])))
            return false;
        void *st = (char *) p->up_ring + (size_t) slot * PLB_UP_SLOT_BYTES;
        memcpy(st, src, size);
        return small_copy(gpu, dst, st, size) &&
               CK(gpu, cudaEventRecord((cudaEvent_t) p->up_event[slot], STREAM(gpu)));
    }
    void *st = plb_gpu_staging(gpu, size);
    if (!st)
        return false;
    // the staging area may still be in flight from a previous upload
    if (!CK(gpu, cudaStreamSynchronize(STREAM(gpu))))
        return false;
    memcpy(st, src, size);
    return CK(gpu, cudaMemcpyAsync(dst, st, size, cudaMemcpyHostToDevice, STREAM(gpu)));
}

bool plb_dev_download(pl_gpu gpu, void *dst, const void *src, size_t size)
{
    DEV_GUARD(gpu);
    void *st = plb_gpu_staging(gpu, size);
    if (!st)
        return false;
    if (size <= PLB_SMALL_XFER) {
        if (!small_copy(gpu, st, src, size))
            return false;
    } else if (!CK(gpu, cudaMemcpyAsync(st, src, size, cudaMemcpyDeviceToHost, STREAM(gpu)))) {
        return false;
    }
    if (!CK(gpu, cudaStreamSynchronize(STREAM(gpu))))
        return false;
    memcpy(dst, st, size);
    return true;
}

// device -> device, asynchronous on the gpu's stream
bool plb_dev_copy(pl_gpu gpu, void *dst, const void *src, size_t size)
{
    DEV_GUARD(gpu);
    if (size <= PLB_SMALL_XFER)
        return small_copy(gpu, dst, src, size);
    return CK(gpu, cudaMemcpyAsync(dst, src, size, cudaMemcpyDeviceToDevice, STREAM(gpu)));
}

// Copy a small payload between device memory and/or pinned host memory with a kernel on
// `stream` (NULL: the gpu's stream) — no DMA engine involved, see k_copy_bytes.
bool pl_cuda_small_copy(pl_gpu gpu, void *dst, const void *src, size_t size, void *stream)
{
    DEV_GUARD(gpu);
    if (!dst || !src || size > PLB_SMALL_XFER) {
        PLB_ERR(gpu->log, "pl_cuda_small_copy: invalid arguments (%zu bytes)", size);
        return false;
    }
    const int blocks = (int) PLB_MIN((size / 16 + 255) / 256 + 1, (size_t) 64);
    k_copy_bytes<<<blocks, 256, 0, stream ? (cudaStream_t) stream : STREAM(gpu)>>>(
        (unsigned char *) dst, (const unsigned char *) src, size);
    return plb_cuda_check(gpu, cudaGetLastError(), "k_copy_bytes");
}

bool plb_dev_memset(pl_gpu gpu, void *dst, int val, size_t size)
{
    DEV_GUARD(gpu);
    return CK(gpu, cudaMemsetAsync(dst, val, size, STREAM(gpu)));
}

// ---------------------------------------------------------------------------------------
// formats — names as in src/vulkan/formats.c:96-120 / src/dummy.c
// ---------------------------------------------------------------------------------------

struct fmt_def {
    const char *name;
    enum pl_fmt_type type;
    int comps, depth, bytes;
    int kind;
};

static const struct fmt_def fmt_defs[] = {
    {"r8",       PL_FMT_UNORM, 1,  8, 1, PLB_FK_UNORM8},
    {"rg8",      PL_FMT_UNORM, 2,  8, 1, PLB_FK_UNORM8},
    {"rgba8",    PL_FMT_UNORM, 4,  8, 1, PLB_FK_UNORM8},
    {"r16",      PL_FMT_UNORM, 1, 16, 2, PLB_FK_UNORM16},
    {"rg16",     PL_FMT_UNORM, 2, 16, 2, PLB_FK_UNORM16},
    {"rgba16",   PL_FMT_UNORM, 4, 16, 2, PLB_FK_UNORM16},
    {"r16hf",    PL_FMT_FLOAT, 1, 16, 2, PLB_FK_HALF},
    {"rg16hf",   PL_FMT_FLOAT, 2, 16, 2, PLB_FK_HALF},
    {"rgba16hf", PL_FMT_FLOAT, 4, 16, 2, PLB_FK_HALF},
    {"r32f",     PL_FMT_FLOAT, 1, 32, 4, PLB_FK_FLOAT},
    {"rg32f",    PL_FMT_FLOAT, 2, 32, 4, PLB_FK_FLOAT},
    {"rgba32f",  PL_FMT_FLOAT, 4, 32, 4, PLB_FK_FLOAT},
    {"r32u",     PL_FMT_UINT,  1, 32, 4, PLB_FK_UINT32},
};

static void setup_formats(struct plb_gpu *p)
{
    const int n = PLB_ARRAY_SIZE(fmt_defs);
    p->fmts = (struct plb_fmt *) calloc(n, sizeof(struct plb_fmt));
    p->fmt_list = (pl_fmt *) calloc(n, sizeof(pl_fmt));
    for (int i = 0; i < n; i++) {
        const struct fmt_def *d = &fmt_defs[i];
        struct pl_fmt_t *f = &p->fmts[i].pub;
        f->name = d->name;
        f->signature = 0xB200F000ull + i;
        f->type = d->type;
        int caps = PL_FMT_CAP_SAMPLEABLE | PL_FMT_CAP_STORABLE | PL_FMT_CAP_RENDERABLE |
                   PL_FMT_CAP_BLITTABLE | PL_FMT_CAP_HOST_READABLE | PL_FMT_CAP_READWRITE;
        if (d->type != PL_FMT_UINT)
            caps |= PL_FMT_CAP_LINEAR | PL_FMT_CAP_BLENDABLE;
        f->caps = (enum pl_fmt_caps) caps;
        f->num_components = d->comps;
        for (int c = 0; c < d->comps; c++) {
            f->component_depth[c] = d->depth;
            f->host_bits[c] = d->bytes * 8;
            f->sample_order[c] = c;
        }
        f->texel_size = f->internal_size = (size_t) d->comps * d->bytes;
        f->texel_align = d->bytes;
        f->gatherable = true;
        p->fmts[i].kind = d->kind;
        p->fmt_list[i] = f;
    }
    p->pub.formats = p->fmt_list;
    p->pub.num_formats = n;
}

// Device-less object that only carries the format table and limits: for format queries
// (pl_find_fmt, pl_find_named_fmt, pl_plane_find_fmt) on machines without a GPU, e.g. when
// planning an ingest layout. Everything else fails on it (no stream, no memory).
pl_gpu pl_cuda_format_catalogue(void)
{
    static struct plb_gpu cat;
    if (!cat.pub.num_formats) {
        setup_formats(&cat);
        cat.pub.limits.max_ssbo_size = (size_t) 1 << 30;
        cat.pub.limits.max_tex_2d_dim = 65536;
        cat.pub.limits.buf_transfer = true;
        cat.pub.glsl.version = 450;
        cat.pub.glsl.compute = true;
        cat.pub.glsl.subgroup_size = 32;
    }
    return &cat.pub;
}

pl_fmt pl_find_named_fmt(pl_gpu gpu, const char *name)
{
    if (!name)
        return NULL;
    for (int i = 0; i < gpu->num_formats; i++) {
        if (strcmp(gpu->formats[i]->name, name) == 0)
            return gpu->formats[i];
    }
    return NULL;
}

// src/gpu.c pl_find_fmt: first format (in list order) matching type, component count,
// minimum depth, host bits and caps
pl_fmt pl_find_fmt(pl_gpu gpu, enum pl_fmt_type type, int num_components, int min_depth,
                   int host_bits, enum pl_fmt_caps caps)
{
    for (int i = 0; i < gpu->num_formats; i++) {
        pl_fmt fmt = gpu->formats[i];
        if (fmt->type != type || fmt->num_components != num_components)
            continue;
        if ((fmt->caps & caps) != caps)
            continue;
        if (fmt->component_depth[0] < min_depth)
            continue;
        if (host_bits && fmt->host_bits[0] != host_bits)
            continue;
        return fmt;
    }
    return NULL;
}

// ---------------------------------------------------------------------------------------
// pl_cuda_create / destroy
// ---------------------------------------------------------------------------------------

pl_gpu pl_cuda_create(pl_log log, const struct pl_cuda_params *params)
{
    static const struct pl_cuda_params defaults = {0};
    params = params ? params : &defaults;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || params->device >= count) {
        // no CPU fallback on this backend: fail loudly
        plb_msg(log, PL_LOG_FATAL, "pl_cuda_create: no usable CUDA device %d (%s)",
                params->device, err != cudaSuccess ? cudaGetErrorString(err) : "out of range");
        return NULL;
    }

    struct plb_gpu *p = (struct plb_gpu *) calloc(1, sizeof(*p));
    p->pub.log = log;
    p->device = params->device;
    p->no_timers = params->no_timers;
    p->async_xfer = !params->no_async_transfers;
    p->coord_mode = params->literal_coords ? PLB_COORD_LITERAL : PLB_COORD_EXACT;
    pl_gpu gpu = &p->pub;
    if (!CK(gpu, cudaSetDevice(p->device)))
        goto error;
    if (params->stream) {
        p->stream = params->stream;
    } else {
        cudaStream_t s;
        if (!CK(gpu, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)))
            goto error;
        p->stream = s;
        p->own_stream = true;
    }
    if (p->async_xfer) {
        cudaStream_t a, b;
        if (!CK(gpu, cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking)) ||
            !CK(gpu, cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking)))
            goto error;
        p->h2d_stream = a;
        p->d2h_stream = b;
        for (int i = 0; i < 64; i++) {
            cudaEvent_t e;
            if (!CK(gpu, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)))
                goto error;
            p->use_ring[i] = e;
        }
    }

    {
        cudaDeviceProp prop;
        if (!CK(gpu, cudaGetDeviceProperties(&prop, p->device)))
            goto error;
        p->sm_count = prop.multiProcessorCount;
        p->pub.glsl.version = 450;
        p->pub.glsl.compute = true;
        p->pub.glsl.max_shmem_size = prop.sharedMemPerBlockOptin;
        p->pub.glsl.max_group_threads = 1024;
        p->pub.glsl.max_group_size[0] = p->pub.glsl.max_group_size[1] = 1024;
        p->pub.glsl.max_group_size[2] = 64;
        p->pub.glsl.subgroup_size = 32;
        p->pub.glsl.min_gather_offset = -32;
        p->pub.glsl.max_gather_offset = 31;
        p->pub.limits.thread_safe = false;
        p->pub.limits.compute_queues = 1; // the stream runs compute (reference: > 0 excludes dummy GPUs)
        p->pub.limits.callbacks = true;
        p->pub.limits.max_buf_size = prop.totalGlobalMem;
        p->pub.limits.max_ssbo_size = prop.totalGlobalMem;
        p->pub.limits.max_tex_1d_dim = 1u << 27;
        p->pub.limits.max_tex_2d_dim = 65536;
        p->pub.limits.max_tex_3d_dim = 4096;
        p->pub.limits.buf_transfer = true;
        p->pub.limits.align_tex_xfer_pitch = 256;
        p->pub.limits.align_tex_xfer_offset = 256;
        p->pub.limits.max_dispatch[0] = 2147483647u;
        p->pub.limits.max_dispatch[1] = p->pub.limits.max_dispatch[2] = 65535;
        PLB_INFO(log, "pl_cuda: device %d '%s' sm_%d%d, %d SMs, %.1f GiB", p->device,
                 prop.name, prop.major, prop.minor, prop.multiProcessorCount,
                 prop.totalGlobalMem / 1073741824.0);
    }

    setup_formats(p);
    if (!plb_upload_pq_consts(gpu))
        goto error;
    return gpu;

error:
    pl_cuda_destroy(&gpu);
    return NULL;
}

static void bank_release(int device);

void pl_cuda_destroy(pl_gpu *pgpu)
{
    if (!pgpu || !*pgpu)
        return;
    struct plb_gpu *p = plb_gpu_priv(*pgpu);
    cudaSetDevice(p->device);
    if (p->stream)
        cudaStreamSynchronize((cudaStream_t) p->stream);
    if (p->bank_ref)
        bank_release(p->device);
    if (p->h2d_stream) {
        cudaStreamSynchronize((cudaStream_t) p->h2d_stream);
        cudaStreamDestroy((cudaStream_t) p->h2d_stream);
    }
    if (p->d2h_stream) {
        cudaStreamSynchronize((cudaStream_t) p->d2h_stream);
        cudaStreamDestroy((cudaStream_t) p->d2h_stream);
    }
    for (int i = 0; i < 64; i++) {
        if (p->use_ring[i])
            cudaEventDestroy((cudaEvent_t) p->use_ring[i]);
    }
    if (p->staging)
        cudaFreeHost(p->staging);
    if (p->up_ring) {
        cudaFreeHost(p->up_ring);
        for (int i = 0; i < PLB_UP_SLOTS; i++) {
            if (p->up_event[i])
                cudaEventDestroy((cudaEvent_t) p->up_event[i]);
        }
    }
    if (p->own_stream && p->stream)
        cudaStreamDestroy((cudaStream_t) p->stream);
    free(p->fmts);
    free(p->fmt_list);
    free(p);
    *pgpu = NULL;
}

void *pl_cuda_stream(pl_gpu gpu) { return plb_gpu_priv(gpu)->stream; }
int plb_gpu_coord_mode(pl_gpu gpu) { return plb_gpu_priv(gpu)->coord_mode; }
bool plb_gpu_no_timers(pl_gpu gpu) { return plb_gpu_priv(gpu)->no_timers; }

bool plb_pq_upload_polar(const plb_pq_consts *q);
bool plb_pq_upload_pixel(const plb_pq_consts *q);
bool plb_pq_upload_ortho(const plb_pq_consts *q);
bool plb_pq_upload_up2a(const plb_pq_consts *q);
bool plb_pq_upload_deint(const plb_pq_consts *q);
bool plb_pq_upload_up2b(const plb_pq_consts *q);
bool plb_pq_upload_up2c(const plb_pq_consts *q);

// Transfer-curve tables (curves.h) live once per device, shared by every pl_gpu on it: the
// descriptors sit in __constant__ memory, which is per device, not per pl_gpu.
#include <pthread.h>
#define PLB_MAX_DEVICES 64
static struct {
    int refs;
    void *tables;
    bool exact;
} g_bank[PLB_MAX_DEVICES];
static pthread_mutex_t g_bank_lock = PTHREAD_MUTEX_INITIALIZER;

static void bank_release(int device)
{
    pthread_mutex_lock(&g_bank_lock);
    if (g_bank[device].refs > 0 && --g_bank[device].refs == 0) {
        cudaFree(g_bank[device].tables);
        g_bank[device].tables = NULL;
    }
    pthread_mutex_unlock(&g_bank_lock);
}

bool plb_upload_pq_consts(pl_gpu gpu)
{
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (p->device < 0 || p->device >= PLB_MAX_DEVICES)
        return false;
    pthread_mutex_lock(&g_bank_lock);
    bool ok = true;
    if (g_bank[p->device].refs == 0) {
        struct plb_pq_consts q;
        plb_pq_literal(&q);
        const char *e = getenv("PLB_EXACT_MATH");
        q.exact_math = e && atoi(e) != 0;
        const size_t total = (size_t) plb_cv_offset(PLB_CV_COUNT) * 4 * sizeof(float);
        float *host = (float *) calloc(1, total);
        void *d = NULL;
        ok = host && cudaMalloc(&d, total) == cudaSuccess;
        for (int i = 0; i < PLB_CV_COUNT && ok; i++) {
            struct plb_curve_host h;
            ok = plb_curve_generate(i, &q, &h);
            if (!ok)
                break;
            memcpy(host + (size_t) plb_cv_offset(i) * 4, h.coef, (size_t) PLB_CV_ENTRIES(i) * 4 * sizeof(float));
            free(h.coef);
            q.cv[i].coef = (const char *) d + (size_t) plb_cv_offset(i) * 16;
            q.cv[i].base = h.base;
            q.cv[i].nseg = q.exact_math ? -1 : h.nseg; // -1: even the segment below the domain misses
        }
        q.cv_bank = d;
        ok = ok && cudaMemcpy(d, host, total, cudaMemcpyHostToDevice) == cudaSuccess &&
             plb_pq_upload_polar(&q) && plb_pq_upload_pixel(&q) && plb_pq_upload_ortho(&q) &&
             plb_pq_upload_up2a(&q) && plb_pq_upload_up2b(&q) && plb_pq_upload_up2c(&q) &&
             plb_pq_upload_deint(&q);
        free(host);
        if (ok) {
            g_bank[p->device].tables = d;
            g_bank[p->device].exact = q.exact_math;
        } else {
            cudaFree(d);
        }
    }
    if (ok) {
        g_bank[p->device].refs++;
        p->bank_ref = true;
        p->exact_math = g_bank[p->device].exact;
    }
    pthread_mutex_unlock(&g_bank_lock);
    if (!ok)
        PLB_ERR(gpu->log, "failed uploading constant tables");
    return ok;
}
uint64_t pl_cuda_launch_count(pl_gpu gpu) { return plb_gpu_priv(gpu)->launches; }

void pl_gpu_flush(pl_gpu gpu) { (void) gpu; }

void pl_gpu_finish(pl_gpu gpu)
{
    DEV_GUARD(gpu);
    struct plb_gpu *p = plb_gpu_priv(gpu);
    CK(gpu, cudaStreamSynchronize(STREAM(gpu)));
    if (p->async_xfer) {
        CK(gpu, cudaStreamSynchronize((cudaStream_t) p->h2d_stream));
        CK(gpu, cudaStreamSynchronize((cudaStream_t) p->d2h_stream));
    }
}

// ---- cross-stream ordering of textures ---------------------------------------------------
// Every main-stream operation that touches textures is numbered; an event is recorded after it
// in a ring of 64. A transfer on a copy stream first waits for the event of the texture's last
// use (or, if that slot was recycled, for the newest event: later, hence still correct).
void plb_gpu_before_use(pl_gpu gpu, pl_tex const *texs, int n)
{
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (!p->async_xfer)
        return;
    for (int i = 0; i < n; i++) {
        struct plb_tex *t = (struct plb_tex *) texs[i];
        if (t && t->parents[0])
            plb_gpu_before_use(gpu, t->parents, 2);
        if (t && t->xfer_pending) {
            CK(gpu, cudaStreamWaitEvent(STREAM(gpu), (cudaEvent_t) t->ev_xfer, 0));
            t->xfer_pending = false;
        }
    }
}

void plb_gpu_after_use(pl_gpu gpu, pl_tex const *texs, int n)
{
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (!p->async_xfer)
        return;
    p->seq++;
    CK(gpu, cudaEventRecord((cudaEvent_t) p->use_ring[p->seq % 64], STREAM(gpu)));
    for (int i = 0; i < n; i++) {
        struct plb_tex *t = (struct plb_tex *) texs[i];
        if (t)
            t->last_use = p->seq;
        for (int k = 0; t && k < 2; k++) {
            if (t->parents[k])
                ((struct plb_tex *) t->parents[k])->last_use = p->seq;
        }
    }
}

static int fmt_kind(pl_fmt fmt);

// Read-only view presenting two single-component textures as one 2-component texture
pl_tex plb_tex_view2(pl_gpu gpu, pl_tex a, pl_tex b, pl_fmt fmt)
{
    const struct plb_tex *ta = plb_tex_priv(a), *tb = plb_tex_priv(b);
    if (ta->d.comps != 1 || tb->d.comps != 1 || ta->d.kind != tb->d.kind || ta->d.ptr2 ||
        tb->d.ptr2 || ta->d.w != tb->d.w || ta->d.h != tb->d.h || fmt->num_components != 2 ||
        fmt_kind(fmt) != ta->d.kind)
        return NULL;
    struct plb_tex *t = (struct plb_tex *) calloc(1, sizeof(*t));
    if (!t)
        return NULL;
    t->pub.params = a->params;
    t->pub.params.format = fmt;
    t->pub.params.host_writable = t->pub.params.host_readable = false;
    t->pub.params.renderable = t->pub.params.storable = t->pub.params.blit_dst = false;
    t->pub.params.sampleable = true;
    t->gpu = gpu;
    t->d = ta->d;
    t->d.comps = 2;
    t->d.ptr2 = tb->d.ptr;
    t->d.pitch2 = tb->d.pitch;
    t->owned = false;
    t->parents[0] = a;
    t->parents[1] = b;
    return &t->pub;
}

// stream for a host<->device transfer of `t`, ordered after everything that used it so far
static cudaStream_t xfer_begin(pl_gpu gpu, struct plb_tex *t, bool upload)
{
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (!p->async_xfer)
        return STREAM(gpu);
    cudaStream_t s = (cudaStream_t) (upload ? p->h2d_stream : p->d2h_stream);
    if (t->last_use) {
        const uint64_t slot = p->seq - t->last_use < 64 ? t->last_use : p->seq;
        CK(gpu, cudaStreamWaitEvent(s, (cudaEvent_t) p->use_ring[slot % 64], 0));
    }
    if (t->xfer_pending) // an earlier transfer on the other copy stream
        CK(gpu, cudaStreamWaitEvent(s, (cudaEvent_t) t->ev_xfer, 0));
    return s;
}

static bool xfer_end(pl_gpu gpu, struct plb_tex *t, cudaStream_t s)
{
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (!p->async_xfer)
        return true;
    if (!t->ev_xfer) {
        cudaEvent_t e;
        if (!CK(gpu, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)))
            return false;
        t->ev_xfer = e;
    }
    t->xfer_pending = true;
    return CK(gpu, cudaEventRecord((cudaEvent_t) t->ev_xfer, s));
}

void pl_cuda_join(pl_gpu gpu)
{
    struct plb_gpu *p = plb_gpu_priv(gpu);
    if (!p->async_xfer)
        return;
    cudaEvent_t tmp[2];
    cudaStream_t s[2] = { (cudaStream_t) p->h2d_stream, (cudaStream_t) p->d2h_stream };
    for (int i = 0; i < 2; i++) {
        if (!CK(gpu, cudaEventCreateWithFlags(&tmp[i], cudaEventDisableTiming)))
            return;
        CK(gpu, cudaEventRecord(tmp[i], s[i]));
        CK(gpu, cudaStreamWaitEvent(STREAM(gpu), tmp[i], 0));
        CK(gpu, cudaEventDestroy(tmp[i])); // released once the event has completed
    }
}

bool pl_gpu_is_failed(pl_gpu gpu) { return plb_gpu_priv(gpu)->failed; }

// ---------------------------------------------------------------------------------------
// textures
// ---------------------------------------------------------------------------------------

static int fmt_kind(pl_fmt fmt) { return ((const struct plb_fmt *) fmt)->kind; }

static bool tex_params_valid(pl_gpu gpu, const struct pl_tex_params *params)
{
    if (!params || !params->format) {
        PLB_ERR(gpu->log, "pl_tex_create: missing params/format");
        return false;
    }
    if (params->w <= 0 || params->h < 0 || params->d < 0) {
        PLB_ERR(gpu->log, "pl_tex_create: invalid dimensions %dx%dx%d", params->w, params->h,
                params->d);
        return false;
    }
    if ((uint32_t) params->w > gpu->limits.max_tex_2d_dim ||
        (uint32_t) params->h > gpu->limits.max_tex_2d_dim) {
        PLB_ERR(gpu->log, "pl_tex_create: dimensions exceed limits");
        return false;
    }
    return true;
}

pl_tex pl_tex_create(pl_gpu gpu, const struct pl_tex_params *params)
{
    DEV_GUARD(gpu);
    if (!tex_params_valid(gpu, params))
        return NULL;
    if (params->export_handle || params->import_handle) {
        PLB_ERR(gpu->log, "pl_tex_create: external memory handles are not supported on the CUDA "
                "backend; wrap device memory with pl_cuda_wrap() instead");
        return NULL;
    }
    struct plb_tex *t = (struct plb_tex *) calloc(1, sizeof(*t));
    t->pub.params = *params;
    t->pub.params.initial_data = NULL;
    t->gpu = gpu;
    const int h = PLB_MAX(params->h, 1), d = PLB_MAX(params->d, 1);
    const size_t row = params->w * params->format->texel_size;
    t->d.pitch = (row + 255) & ~(size_t) 255;
    t->d.w = params->w;
    t->d.h = h * d; // 3D textures are stored as d slices of h rows
    t->d.comps = params->format->num_components;
    t->d.kind = fmt_kind(params->format);
    t->alloc_size = t->d.pitch * h * d;
    t->owned = true;
    if (!CK(gpu, cudaMalloc(&t->d.ptr, t->alloc_size))) {
        free(t);
        return NULL;
    }
    if (params->initial_data) {
        struct pl_tex_transfer_params tp = {0};
        tp.tex = &t->pub;
        tp.ptr = (void *) params->initial_data;
        if (!pl_tex_upload(gpu, &tp)) {
            pl_tex tex = &t->pub;
            pl_tex_destroy(gpu, &tex);
            return NULL;
        }
    }
    return &t->pub;
}

pl_tex pl_cuda_wrap(pl_gpu gpu, const struct pl_cuda_wrap_params *params)
{
    if (!params || !params->ptr || !params->format || params->w <= 0 || params->h <= 0) {
        PLB_ERR(gpu->log, "pl_cuda_wrap: invalid parameters");
        return NULL;
    }
    const size_t row = params->w * params->format->texel_size;
    if (params->pitch < row || params->pitch % params->format->texel_align) {
        PLB_ERR(gpu->log, "pl_cuda_wrap: pitch %zu too small/misaligned for %d texels of %s",
                params->pitch, params->w, params->format->name);
        return NULL;
    }
    // The kernels access texels as whole vectors (uchar4 / ushort4 / uint2 / float4) and store
    // runs of four texels as 16-byte words: the base address and the pitch must be multiples
    // of 16 bytes (cudaMalloc / cudaMallocPitch / NVDEC surfaces are 256-byte aligned).
    if (((uintptr_t) params->ptr & 15) || (params->pitch & 15)) {
        PLB_ERR(gpu->log, "pl_cuda_wrap: pointer %p / pitch %zu must be 16-byte aligned",
                params->ptr, params->pitch);
        return NULL;
    }
    struct plb_tex *t = (struct plb_tex *) calloc(1, sizeof(*t));
    t->pub.params.w = params->w;
    t->pub.params.h = params->h;
    t->pub.params.format = params->format;
    t->pub.params.sampleable = t->pub.params.renderable = t->pub.params.storable = true;
    t->pub.params.host_writable = t->pub.params.host_readable = true;
    t->pub.params.blit_src = t->pub.params.blit_dst = true;
    t->gpu = gpu;
    t->d.ptr = params->ptr;
    t->d.pitch = params->pitch;
    t->d.w = params->w;
    t->d.h = params->h;
    t->d.comps = params->format->num_components;
    t->d.kind = fmt_kind(params->format);
    t->owned = false;
    return &t->pub;
}

void *pl_cuda_tex_ptr(pl_tex tex, size_t *out_pitch)
{
    const struct plb_tex *t = plb_tex_priv(tex);
    if (out_pitch)
        *out_pitch = t->d.pitch;
    return t->d.ptr;
}

void pl_tex_destroy(pl_gpu gpu, pl_tex *ptex)
{
    DEV_GUARD(gpu);
    if (!ptex || !*ptex)
        return;
    struct plb_tex *t = (struct plb_tex *) *ptex;
    if (t->xfer_pending)
        CK(gpu, cudaStreamWaitEvent(STREAM(gpu), (cudaEvent_t) t->ev_xfer, 0));
    if (t->owned && t->d.ptr)
        CK(gpu, cudaFreeAsync(t->d.ptr, STREAM(gpu)));
    if (t->ev_xfer)
        cudaEventDestroy((cudaEvent_t) t->ev_xfer);
    free(t);
    *ptex = NULL;
}

// src/gpu.c:472-520: keep the texture if the parameters are compatible
bool pl_tex_recreate(pl_gpu gpu, pl_tex *tex, const struct pl_tex_params *params)
{
    if (params->initial_data) {
        PLB_ERR(gpu->log, "pl_tex_recreate may not be used with `initial_data`!");
        return false;
    }
    if (*tex) {
        // pl_tex_params_superset (src/gpu.c:276-304): same size and format, and every usage the
        // caller asks for is already available
        const struct pl_tex_params *cur = &(*tex)->params;
        const bool usage_ok = cur->sampleable >= params->sampleable && cur->renderable >= params->renderable &&
                              cur->storable >= params->storable && cur->blit_src >= params->blit_src &&
                              cur->blit_dst >= params->blit_dst && cur->host_writable >= params->host_writable &&
                              cur->host_readable >= params->host_readable;
        if (cur->w == params->w && cur->h == params->h && cur->d == params->d &&
            cur->format == params->format && usage_ok)
            return true;
        pl_tex_destroy(gpu, tex);
    }
    *tex = pl_tex_create(gpu, params);
    return !!*tex;
}

void pl_tex_invalidate(pl_gpu gpu, pl_tex tex) { (void) gpu; (void) tex; }

extern bool plb_launch_clear(pl_gpu gpu, const struct plb_tex_desc *dst, const float color[4]);

void pl_tex_clear(pl_gpu gpu, pl_tex dst, const float color[4])
{
    DEV_GUARD(gpu);
    plb_gpu_before_use(gpu, &dst, 1);
    plb_launch_clear(gpu, &plb_tex_priv(dst)->d, color);
    plb_gpu_after_use(gpu, &dst, 1);
}

} // extern "C"

// 32-bit words of an integer texture <- the clear value of their component
__global__ void k_fill_words(char *ptr, size_t pitch, int words, int rows, int comps, uint4 v)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= words || y >= rows)
        return;
    const int c = x % comps;
    ((uint32_t *) (ptr + (size_t) y * pitch))[x] = c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}

extern "C" {

void pl_tex_clear_ex(pl_gpu gpu, pl_tex dst, const union pl_clear_color color)
{
    const enum pl_fmt_type type = dst->params.format->type;
    if (type == PL_FMT_UINT || type == PL_FMT_SINT) {
        // integer texels take the value itself (gpu.h:792-806); 32-bit components
        DEV_GUARD(gpu);
        pl_fmt fmt = dst->params.format;
        const struct plb_tex *t = plb_tex_priv(dst);
        if (fmt->texel_size != 4u * fmt->num_components) {
            PLB_ERR(gpu->log, "pl_tex_clear_ex: unsupported integer format %s", fmt->name);
            return;
        }
        plb_gpu_before_use(gpu, &dst, 1);
        const int words = t->d.w * fmt->num_components;
        dim3 block(64, 4), grid((words + 63) / 64, (t->d.h + 3) / 4);
        k_fill_words<<<grid, block, 0, STREAM(gpu)>>>((char *) t->d.ptr, t->d.pitch, words, t->d.h,
                                                      fmt->num_components,
                                                      make_uint4(color.u[0], color.u[1], color.u[2], color.u[3]));
        plb_gpu_priv(gpu)->launches++;
        plb_cuda_check(gpu, cudaGetLastError(), "k_fill_words");
        plb_gpu_after_use(gpu, &dst, 1);
        return;
    }
    pl_tex_clear(gpu, dst, color.f);
}

// pl_tex_blit (gpu.h:820-849): copy / scale a sub-rectangle between textures of loosely
// compatible formats. Lowered to one pixel-pass launch: the source rectangle is sampled (texel
// exact for equal sizes, nearest / bilinear otherwise) onto the destination rectangle; flipped
// rectangles sample in reverse. Everything outside dst_rc is preserved.
void pl_tex_blit(pl_gpu gpu, const struct pl_tex_blit_params *params)
{
    DEV_GUARD(gpu);
    pl_tex src = params ? params->src : NULL, dst = params ? params->dst : NULL;
    if (!src || !dst) {
        PLB_ERR(gpu->log, "pl_tex_blit: src and dst are required");
        return;
    }
    pl_fmt sf = src->params.format, df = dst->params.format;
    if (sf->internal_size != df->internal_size || (sf->type == PL_FMT_UINT) != (df->type == PL_FMT_UINT) ||
        (sf->type == PL_FMT_SINT) != (df->type == PL_FMT_SINT) || !src->params.blit_src ||
        !dst->params.blit_dst ||
        (params->sample_mode == PL_TEX_SAMPLE_LINEAR && !(sf->caps & PL_FMT_CAP_LINEAR))) {
        PLB_ERR(gpu->log, "pl_tex_blit: incompatible textures (%s -> %s) or missing blit usage",
                sf->name, df->name);
        return;
    }
    if (src->params.d || dst->params.d) {
        PLB_ERR(gpu->log, "pl_tex_blit: 3D textures are not blittable on the CUDA backend");
        return;
    }
    pl_rect3d sr = params->src_rc, dr = params->dst_rc;
    if (!sr.x0 && !sr.x1) sr.x1 = src->params.w;
    if (!sr.y0 && !sr.y1) sr.y1 = PLB_MAX(src->params.h, 1);
    if (!dr.x0 && !dr.x1) dr.x1 = dst->params.w;
    if (!dr.y0 && !dr.y1) dr.y1 = PLB_MAX(dst->params.h, 1);
    // same format, same size, upright: a plain copy of the texels' bytes (any format, any bits)
    if (sf == df && sr.x1 - sr.x0 == dr.x1 - dr.x0 && sr.y1 - sr.y0 == dr.y1 - dr.y0 &&
        sr.x1 > sr.x0 && sr.y1 > sr.y0 && dr.x1 > dr.x0 && dr.y1 > dr.y0 && sr.x0 >= 0 && sr.y0 >= 0 &&
        dr.x0 >= 0 && dr.y0 >= 0 && sr.x1 <= src->params.w && sr.y1 <= PLB_MAX(src->params.h, 1) &&
        dr.x1 <= dst->params.w && dr.y1 <= PLB_MAX(dst->params.h, 1)) {
        const struct plb_tex *st = plb_tex_priv(src), *dt = plb_tex_priv(dst);
        const size_t ts = sf->texel_size;
        pl_tex texs[2] = { src, dst };
        plb_gpu_before_use(gpu, texs, 2);
        CK(gpu, cudaMemcpy2DAsync((char *) dt->d.ptr + (size_t) dr.y0 * dt->d.pitch + dr.x0 * ts, dt->d.pitch,
                                  (const char *) st->d.ptr + (size_t) sr.y0 * st->d.pitch + sr.x0 * ts, st->d.pitch,
                                  (size_t) (sr.x1 - sr.x0) * ts, (size_t) (sr.y1 - sr.y0),
                                  cudaMemcpyDeviceToDevice, STREAM(gpu)));
        plb_gpu_after_use(gpu, texs, 2);
        return;
    }
    if (sf->type == PL_FMT_UINT || sf->type == PL_FMT_SINT) {
        PLB_ERR(gpu->log, "pl_tex_blit: scaled / converting blits of integer textures are not "
                "supported on the CUDA backend");
        return;
    }
    const bool in_src = PLB_MIN(sr.x0, sr.x1) >= 0 && PLB_MAX(sr.x0, sr.x1) <= src->params.w &&
                        PLB_MIN(sr.y0, sr.y1) >= 0 && PLB_MAX(sr.y0, sr.y1) <= src->params.h;
    const bool in_dst = PLB_MIN(dr.x0, dr.x1) >= 0 && PLB_MAX(dr.x0, dr.x1) <= dst->params.w &&
                        PLB_MIN(dr.y0, dr.y1) >= 0 && PLB_MAX(dr.y0, dr.y1) <= dst->params.h;
    if (!in_src || !in_dst || sr.x0 == sr.x1 || sr.y0 == sr.y1 || dr.x0 == dr.x1 || dr.y0 == dr.y1) {
        PLB_ERR(gpu->log, "pl_tex_blit: rectangles out of bounds or empty");
        return;
    }
    // a flipped destination rectangle = an upright one sampled in reverse
    if (dr.x0 > dr.x1) { int t = dr.x0; dr.x0 = dr.x1; dr.x1 = t; t = sr.x0; sr.x0 = sr.x1; sr.x1 = t; }
    if (dr.y0 > dr.y1) { int t = dr.y0; dr.y0 = dr.y1; dr.y1 = t; t = sr.y0; sr.y0 = sr.y1; sr.y1 = t; }

    struct plb_pixel_launch *px = (struct plb_pixel_launch *) calloc(1, sizeof(*px));
    struct plb_plane_src *pl = &px->planes[0];
    px->nplanes = 1;
    pl->kind = PLB_PLANE_DIRECT;
    pl->tex = plb_tex_priv(src)->d;
    pl->rect = (pl_rect2df) { (float) sr.x0, (float) sr.y0, (float) sr.x1, (float) sr.y1 };
    const bool same_size = abs(sr.x1 - sr.x0) == dr.x1 - dr.x0 && abs(sr.y1 - sr.y0) == dr.y1 - dr.y0;
    pl->linear = !same_size && params->sample_mode == PL_TEX_SAMPLE_LINEAR;
    pl->scale = 1.0f;
    for (int k = 0; k < 4; k++)
        pl->map[k] = k;
    px->init[3] = 1.0f;
    px->dst = plb_tex_priv(dst)->d;
    px->dst_rect = (pl_rect2d) { dr.x0, dr.y0, dr.x1, dr.y1 };
    px->out_w = dr.x1 - dr.x0;
    px->out_h = dr.y1 - dr.y0;
    px->coord_mode = PLB_COORD_EXACT;
    pl_tex texs[2] = { src, dst };
    plb_gpu_before_use(gpu, texs, 2);
    plb_launch_pixel(gpu, px);
    plb_gpu_after_use(gpu, texs, 2);
    free(px);
}

// Raw texel copy between two textures of the same format (renderer: over-rendered planes)
bool plb_tex_copy_region(pl_gpu gpu, pl_tex dst, int dx, int dy, pl_tex src, int sx, int sy, int w, int h)
{
    DEV_GUARD(gpu);
    if (src->params.format != dst->params.format || w <= 0 || h <= 0 || dx < 0 || dy < 0 || sx < 0 ||
        sy < 0 || dx + w > dst->params.w || dy + h > PLB_MAX(dst->params.h, 1) ||
        sx + w > src->params.w || sy + h > PLB_MAX(src->params.h, 1)) {
        PLB_ERR(gpu->log, "internal texture copy: bad region");
        return false;
    }
    const struct plb_tex *st = plb_tex_priv(src), *dt = plb_tex_priv(dst);
    const size_t ts = src->params.format->texel_size;
    pl_tex texs[2] = { src, dst };
    plb_gpu_before_use(gpu, texs, 2);
    const bool ok = CK(gpu, cudaMemcpy2DAsync((char *) dt->d.ptr + (size_t) dy * dt->d.pitch + dx * ts, dt->d.pitch,
                                              (const char *) st->d.ptr + (size_t) sy * st->d.pitch + sx * ts,
                                              st->d.pitch, (size_t) w * ts, (size_t) h,
                                              cudaMemcpyDeviceToDevice, STREAM(gpu)));
    plb_gpu_after_use(gpu, texs, 2);
    return ok;
}

bool pl_fmt_is_ordered(pl_fmt fmt)
{
    bool ordered = true;
    for (int i = 0; i < fmt->num_components; i++)
        ordered &= fmt->sample_order[i] == i;
    return ordered;
}

bool pl_fmt_is_float(pl_fmt fmt)
{
    return fmt->type != PL_FMT_UINT && fmt->type != PL_FMT_SINT; // unknown counts as float (gpu.c:61-75)
}

static void CUDART_CB host_cb_trampoline(void *data)
{
    void **pair = (void **) data;
    void (*cb)(void *) = (void (*)(void *)) pair[0];
    cb(pair[1]);
    free(pair);
}

static bool fix_transfer(pl_gpu gpu, const struct pl_tex_transfer_params *params,
                         pl_rect3d *rc, size_t *row_pitch, const char *what)
{
    pl_tex tex = params->tex;
    if (!tex) {
        PLB_ERR(gpu->log, "%s: no texture", what);
        return false;
    }
    *rc = params->rc;
    // src/gpu.c strip_coords / infer_rc: empty rect = whole texture
    if (!rc->x0 && !rc->x1) rc->x1 = tex->params.w;
    if (!rc->y0 && !rc->y1) rc->y1 = PLB_MAX(tex->params.h, 1);
    if (tex->params.d) {
        // 3D textures are d slices of h rows: whole-slice transfers with the default depth pitch
        // are runs of rows; anything else is outside what the path needs
        if (!rc->z0 && !rc->z1) rc->z1 = tex->params.d;
        const int h = tex->params.h;
        const size_t rp = params->row_pitch ? params->row_pitch
                        : (size_t) (rc->x1 - rc->x0) * tex->params.format->texel_size;
        if (rc->y0 != 0 || rc->y1 != h || rc->z0 < 0 || rc->z1 > tex->params.d || rc->z1 <= rc->z0 ||
            (params->depth_pitch && params->depth_pitch != rp * h)) {
            PLB_ERR(gpu->log, "%s: only whole slices of a 3D texture can be transferred", what);
            return false;
        }
        rc->y0 = rc->z0 * h;
        rc->y1 = rc->z1 * h;
    }
    if (rc->x0 < 0 || rc->y0 < 0 || rc->x1 > tex->params.w ||
        rc->y1 > PLB_MAX(tex->params.h, 1) * PLB_MAX(tex->params.d, 1) ||
        rc->x1 <= rc->x0 || rc->y1 <= rc->y0) {
        PLB_ERR(gpu->log, "%s: rect out of bounds", what);
        return false;
    }
    size_t min_pitch = (size_t) (rc->x1 - rc->x0) * tex->params.format->texel_size;
    *row_pitch = params->row_pitch ? params->row_pitch : min_pitch;
    if (*row_pitch < min_pitch) {
        PLB_ERR(gpu->log, "%s: row_pitch %zu smaller than a row (%zu)", what, *row_pitch, min_pitch);
        return false;
    }
    if (!params->ptr && !params->buf) {
        PLB_ERR(gpu->log, "%s: neither ptr nor buf given", what);
        return false;
    }
    return true;
}

bool pl_tex_upload(pl_gpu gpu, const struct pl_tex_transfer_params *params)
{
    DEV_GUARD(gpu);
    pl_rect3d rc;
    size_t row_pitch;
    if (!fix_transfer(gpu, params, &rc, &row_pitch, "pl_tex_upload"))
        return false;
    struct plb_tex *t = (struct plb_tex *) params->tex;
    const size_t ts = params->tex->params.format->texel_size;
    char *dst = (char *) t->d.ptr + (size_t) rc.y0 * t->d.pitch + rc.x0 * ts;
    const size_t wbytes = (size_t) (rc.x1 - rc.x0) * ts;
    bool ok;
    cudaStream_t s = STREAM(gpu);
    if (params->buf) {
        const struct plb_buf *b = (const struct plb_buf *) params->buf;
        const char *src = (const char *) (b->dptr ? b->dptr : b->hptr) + params->buf_offset;
        pl_tex one = params->tex;
        plb_gpu_before_use(gpu, &one, 1);
        ok = CK(gpu, cudaMemcpy2DAsync(dst, t->d.pitch, src, row_pitch, wbytes, rc.y1 - rc.y0,
                                       b->dptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                       s));
        plb_gpu_after_use(gpu, &one, 1);
    } else {
        s = xfer_begin(gpu, t, true);
        if (row_pitch == wbytes && t->d.pitch == wbytes) // contiguous on both sides: one linear DMA
            ok = CK(gpu, cudaMemcpyAsync(dst, params->ptr, wbytes * (size_t) (rc.y1 - rc.y0),
                                         cudaMemcpyHostToDevice, s));
        else
            ok = CK(gpu, cudaMemcpy2DAsync(dst, t->d.pitch, params->ptr, row_pitch, wbytes,
                                           rc.y1 - rc.y0, cudaMemcpyHostToDevice, s));
        ok = ok && xfer_end(gpu, t, s);
    }
    if (ok && params->callback) {
        void **pair = (void **) malloc(2 * sizeof(void *));
        pair[0] = (void *) params->callback;
        pair[1] = params->priv;
        ok = CK(gpu, cudaLaunchHostFunc(s, host_cb_trampoline, pair));
    } else if (ok && !params->buf) {
        // gpu.h:885-888: without a callback the caller may reuse or free `ptr` on return. The
        // copy from pinned memory is a DMA still in flight here, so wait for it (pageable
        // memory was staged by the driver already). Pass a callback for asynchronous uploads.
        ok = CK(gpu, cudaStreamSynchronize(s));
    }
    return ok;
}

bool pl_tex_download(pl_gpu gpu, const struct pl_tex_transfer_params *params)
{
    DEV_GUARD(gpu);
    pl_rect3d rc;
    size_t row_pitch;
    if (!fix_transfer(gpu, params, &rc, &row_pitch, "pl_tex_download"))
        return false;
    struct plb_tex *t = (struct plb_tex *) params->tex;
    const size_t ts = params->tex->params.format->texel_size;
    const char *src = (const char *) t->d.ptr + (size_t) rc.y0 * t->d.pitch + rc.x0 * ts;
    const size_t wbytes = (size_t) (rc.x1 - rc.x0) * ts;
    bool ok;
    cudaStream_t s = STREAM(gpu);
    if (params->buf) {
        const struct plb_buf *b = (const struct plb_buf *) params->buf;
        char *dst = (char *) (b->dptr ? b->dptr : b->hptr) + params->buf_offset;
        pl_tex one = params->tex;
        plb_gpu_before_use(gpu, &one, 1);
        ok = CK(gpu, cudaMemcpy2DAsync(dst, row_pitch, src, t->d.pitch, wbytes, rc.y1 - rc.y0,
                                       b->dptr ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                       s));
        plb_gpu_after_use(gpu, &one, 1);
    } else {
        s = xfer_begin(gpu, t, false);
        if (row_pitch == wbytes && t->d.pitch == wbytes)
            ok = CK(gpu, cudaMemcpyAsync(params->ptr, src, wbytes * (size_t) (rc.y1 - rc.y0),
                                         cudaMemcpyDeviceToHost, s));
        else
            ok = CK(gpu, cudaMemcpy2DAsync(params->ptr, row_pitch, src, t->d.pitch, wbytes,
                                           rc.y1 - rc.y0, cudaMemcpyDeviceToHost, s));
        ok = ok && xfer_end(gpu, t, s);
    }
    if (!ok)
        return false;
    if (params->callback) {
        void **pair = (void **) malloc(2 * sizeof(void *));
        pair[0] = (void *) params->callback;
        pair[1] = params->priv;
        return CK(gpu, cudaLaunchHostFunc(s, host_cb_trampoline, pair));
    }
    // blocking semantics when no callback/buffer polling is requested (gpu.h:905-910)
    if (!params->buf)
        return CK(gpu, cudaStreamSynchronize(s));
    return true;
}

bool pl_tex_poll(pl_gpu gpu, pl_tex tex, uint64_t timeout)
{
    DEV_GUARD(gpu);
    (void) tex;
    if (timeout) {
        CK(gpu, cudaStreamSynchronize(STREAM(gpu)));
        return false;
    }
    return cudaStreamQuery(STREAM(gpu)) == cudaErrorNotReady;
}

// ---------------------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------------------

pl_buf pl_buf_create(pl_gpu gpu, const struct pl_buf_params *params)
{
    DEV_GUARD(gpu);
    if (!params || !params->size) {
        PLB_ERR(gpu->log, "pl_buf_create: zero-sized buffer");
        return NULL;
    }
    if (params->export_handle || params->import_handle) {
        PLB_ERR(gpu->log, "pl_buf_create: external memory handles are not supported on the CUDA backend");
        return NULL;
    }
    struct plb_buf *b = (struct plb_buf *) calloc(1, sizeof(*b));
    b->pub.params = *params;
    b->pub.params.initial_data = NULL;
    bool host_only = params->memory_type == PL_BUF_MEM_HOST ||
                     (params->host_mapped && !params->storable && !params->uniform);
    if (!host_only && !CK(gpu, cudaMalloc(&b->dptr, params->size)))
        goto error;
    if (host_only || params->host_mapped || params->host_readable || params->host_writable) {
        if (!CK(gpu, cudaMallocHost(&b->hptr, params->size)))
            goto error;
        memset(b->hptr, 0, params->size);
    }
    if (params->host_mapped)
        b->pub.data = (uint8_t *) b->hptr;
    if (params->initial_data) {
        if (b->hptr)
            memcpy(b->hptr, params->initial_data, params->size);
        if (b->dptr) {
            if (b->hptr) {
                if (!CK(gpu, cudaMemcpyAsync(b->dptr, b->hptr, params->size,
                                             cudaMemcpyHostToDevice, STREAM(gpu))))
                    goto error;
            } else if (!plb_dev_upload(gpu, b->dptr, params->initial_data, params->size)) {
                goto error;
            }
        }
    } else if (b->dptr) {
        CK(gpu, cudaMemsetAsync(b->dptr, 0, params->size, STREAM(gpu)));
    }
    return &b->pub;

error:
    if (b->dptr) cudaFree(b->dptr);
    if (b->hptr) cudaFreeHost(b->hptr);
    free(b);
    return NULL;
}

void pl_buf_destroy(pl_gpu gpu, pl_buf *pbuf)
{
    DEV_GUARD(gpu);
    if (!pbuf || !*pbuf)
        return;
    struct plb_buf *b = (struct plb_buf *) *pbuf;
    if (b->hptr) {
        CK(gpu, cudaStreamSynchronize(STREAM(gpu)));
        cudaFreeHost(b->hptr);
    }
    if (b->dptr)
        CK(gpu, cudaFreeAsync(b->dptr, STREAM(gpu)));
    free(b);
    *pbuf = NULL;
}

bool pl_buf_recreate(pl_gpu gpu, pl_buf *buf, const struct pl_buf_params *params)
{
    if (*buf && (*buf)->params.size >= params->size)
        return true;
    pl_buf_destroy(gpu, buf);
    *buf = pl_buf_create(gpu, params);
    return !!*buf;
}

void pl_buf_write(pl_gpu gpu, pl_buf buf, size_t buf_offset, const void *data, size_t size)
{
    DEV_GUARD(gpu);
    struct plb_buf *b = (struct plb_buf *) buf;
    if (buf_offset + size > buf->params.size) {
        PLB_ERR(gpu->log, "pl_buf_write: out of bounds");
        return;
    }
    if (b->hptr) {
        // host mirror may be read by in-flight copies
        CK(gpu, cudaStreamSynchronize(STREAM(gpu)));
        memcpy((char *) b->hptr + buf_offset, data, size);
        if (b->dptr)
            CK(gpu, cudaMemcpyAsync((char *) b->dptr + buf_offset, (char *) b->hptr + buf_offset,
                                    size, cudaMemcpyHostToDevice, STREAM(gpu)));
    } else {
        plb_dev_upload(gpu, (char *) b->dptr + buf_offset, data, size);
    }
}

bool pl_buf_read(pl_gpu gpu, pl_buf buf, size_t buf_offset, void *dest, size_t size)
{
    DEV_GUARD(gpu);
    struct plb_buf *b = (struct plb_buf *) buf;
    if (buf_offset + size > buf->params.size) {
        PLB_ERR(gpu->log, "pl_buf_read: out of bounds");
        return false;
    }
    if (b->dptr)
        return plb_dev_download(gpu, dest, (char *) b->dptr + buf_offset, size);
    if (!CK(gpu, cudaStreamSynchronize(STREAM(gpu))))
        return false;
    memcpy(dest, (char *) b->hptr + buf_offset, size);
    return true;
}

void pl_buf_copy(pl_gpu gpu, pl_buf dst, size_t dst_offset, pl_buf src, size_t src_offset,
                 size_t size)
{
    DEV_GUARD(gpu);
    struct plb_buf *d = (struct plb_buf *) dst, *s = (struct plb_buf *) src;
    void *dp = (char *) (d->dptr ? d->dptr : d->hptr) + dst_offset;
    const void *sp = (const char *) (s->dptr ? s->dptr : s->hptr) + src_offset;
    CK(gpu, cudaMemcpyAsync(dp, sp, size, cudaMemcpyDefault, STREAM(gpu)));
}

// pl_buf_copy_swap (src/gpu.h:149-165, the reference runs it as a compute shader over SSBOs)
__global__ void k_swap_words(const uint32_t *src, uint32_t *dst, size_t nwords, int wordsize)
{
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nwords;
         i += (size_t) gridDim.x * blockDim.x) {
        const uint32_t v = src[i];
        // 16-bit words: swap the bytes of each half; 32-bit words: reverse all four
        dst[i] = wordsize == 2 ? __byte_perm(v, 0, 0x2301) : __byte_perm(v, 0, 0x0123);
    }
}

bool plb_buf_copy_swap(pl_gpu gpu, pl_buf dst, size_t dst_offset, pl_buf src, size_t src_offset,
                       size_t size, int wordsize)
{
    DEV_GUARD(gpu);
    struct plb_buf *d = (struct plb_buf *) dst, *s = (struct plb_buf *) src;
    if ((wordsize != 2 && wordsize != 4) || size % 4 || dst_offset % 4 || src_offset % 4 ||
        !d->dptr || !s->dptr || src_offset + size > s->pub.params.size ||
        dst_offset + size > d->pub.params.size) {
        PLB_ERR(gpu->log, "pl_buf_copy_swap: invalid arguments (word size %d, %zu bytes)", wordsize, size);
        return false;
    }
    if (!size)
        return true;
    const size_t nwords = size / 4;
    const int grid = (int) PLB_MIN((nwords + 255) / 256, (size_t) (plb_gpu_priv(gpu)->sm_count * 8));
    k_swap_words<<<grid, 256, 0, STREAM(gpu)>>>((const uint32_t *) ((const char *) s->dptr + src_offset),
                                                (uint32_t *) ((char *) d->dptr + dst_offset), nwords, wordsize);
    plb_gpu_priv(gpu)->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_swap_words");
}

// src/gpu.h:145-165 (library-internal in the reference; its conformance test calls it)
bool pl_buf_copy_swap(pl_gpu gpu, const struct pl_buf_copy_swap_params *params)
{
    return plb_buf_copy_swap(gpu, params->dst, params->dst_offset, params->src, params->src_offset,
                             params->size, params->wordsize);
}

bool pl_buf_poll(pl_gpu gpu, pl_buf buf, uint64_t timeout)
{
    DEV_GUARD(gpu);
    return pl_tex_poll(gpu, NULL, timeout) && buf;
}

// ---------------------------------------------------------------------------------------
// timers: cudaEvent pairs recorded around dispatches (reference: gpu.h:415-437)
// ---------------------------------------------------------------------------------------

struct pl_timer_t {
    cudaEvent_t start, stop;
    bool recorded;
};

pl_timer pl_timer_create(pl_gpu gpu)
{
    DEV_GUARD(gpu);
    struct pl_timer_t *t = (struct pl_timer_t *) calloc(1, sizeof(*t));
    if (!CK(gpu, cudaEventCreate(&t->start)) || !CK(gpu, cudaEventCreate(&t->stop))) {
        free(t);
        return NULL;
    }
    return t;
}

void pl_timer_destroy(pl_gpu gpu, pl_timer *pt)
{
    DEV_GUARD(gpu);
    (void) gpu;
    if (!pt || !*pt)
        return;
    cudaEventDestroy((*pt)->start);
    cudaEventDestroy((*pt)->stop);
    free(*pt);
    *pt = NULL;
}

void plb_timer_begin(pl_gpu gpu, pl_timer t)
{
    DEV_GUARD(gpu);
    if (t)
        CK(gpu, cudaEventRecord(t->start, STREAM(gpu)));
}

void plb_timer_end(pl_gpu gpu, pl_timer t)
{
    DEV_GUARD(gpu);
    if (t) {
        CK(gpu, cudaEventRecord(t->stop, STREAM(gpu)));
        t->recorded = true;
    }
}

void plb_timer_wait(pl_gpu gpu, pl_timer t)
{
    DEV_GUARD(gpu);
    if (t && t->recorded)
        CK(gpu, cudaEventSynchronize(t->stop));
}

// nanoseconds of the last recorded interval, 0 if not available yet
uint64_t pl_timer_query(pl_gpu gpu, pl_timer t)
{
    DEV_GUARD(gpu);
    (void) gpu;
    if (!t || !t->recorded)
        return 0;
    if (cudaEventQuery(t->stop) != cudaSuccess)
        return 0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t->start, t->stop) != cudaSuccess)
        return 0;
    t->recorded = false;
    return (uint64_t) (ms * 1e6);
}

} // extern "C"
