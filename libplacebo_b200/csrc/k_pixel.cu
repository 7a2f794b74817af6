// Plain pixel pass: what the reference does with a fragment/compute shader that samples up to
// four planes 1:1 (or through a free bilinear/nearest resize), merges them, runs colour ops and
// writes the target — pass_read_image's merge shader (src/renderer.c:1788-1897), the FBO ->
// colour-map -> dither -> target pass (renderer.c:2157-2279, 2586-2964) and the standalone
// pl_dispatch_finish of sample_direct/deband shaders.
//
// CTA = 32x8 threads covering a 32x32 output tile (4 rows per thread, x fastest so that
// loads/stores of a warp are one contiguous row segment). The tile is a whole number of
// peak-detection work groups (8/16/32 wide), see pixel_core.cuh.
#include "pixel_core.cuh"

PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_pixel)

#define PIX_TILE 32

// MINB = resident CTAs per SM the register allocation is tuned for (2 -> 114 registers, 3 -> 80,
// 4 -> 64, 5 -> 48). Measured on B200 with the HDR colour pass (µs): 1305 / 1186 / 1158 / 1138;
// debanding: 106 / 98 / 93 / 103. Default 4; PLB_PIXEL_MINB selects another for measurements.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_pixel(const __grid_constant__ plb_tail_args ap)
{
    __shared__ plb_peak_smem ps;
    plm_tables_init();
    plb_tail_stage(ap);
    __syncthreads();
    const plb_tail_args &a = plb_s_tail;
    const int X0 = blockIdx.x * PIX_TILE, Y0 = blockIdx.y * PIX_TILE;
    const int gx = X0 + threadIdx.x;

    // Does the program contain a peak tap? (uniform)
    bool has_peak = false;
    for (int i = 0; i < plb_s_prog.n; i++)
        has_peak |= plb_s_prog.ops[i].kind == PLB_OP_PEAK;
    const int wgs_x = has_peak ? PIX_TILE / a.wg_w : 1, wgs_y = has_peak ? PIX_TILE / a.wg_h : 1;
    const int nwgx = has_peak ? (a.out_w + a.wg_w - 1) / a.wg_w : 1;
    const int nwgy = has_peak ? (a.out_h + a.wg_h - 1) / a.wg_h : 1;
    if (has_peak)
        plb_peak_begin(ps);

    const plb_op_peak *pk_any = nullptr;
    #pragma unroll 1
    for (int r = 0; r < PIX_TILE / 8; r++) {
        const int gy = Y0 + threadIdx.y + 8 * r;
        const bool inside = gx < a.out_w && gy < a.out_h;
        if (!has_peak && !inside)
            continue;
        float color[4];
        uint32_t ypq = 0;
        const plb_op_peak *pk = nullptr;
        plb_tail_color(gx, gy, nullptr, color, &ypq, &pk);
        if (has_peak) {
            const int lx = (gx - X0) / a.wg_w, ly = (gy - Y0) / a.wg_h;
            const int wg_local = ly * wgs_x + lx;
            const bool valid = (X0 / a.wg_w + lx) < nwgx && (Y0 / a.wg_h + ly) < nwgy;
            pk_any = pk;
            plb_peak_add(ps, *pk, valid, wg_local, ypq);
        }
        if (inside && a.dst.ptr)
            plb_store(a.dst, a.dst_x0 + gx, a.dst_y0 + gy,
                      make_float4(color[0], color[1], color[2], color[3]));
    }

    if (has_peak)
        plb_peak_finish(ps, *pk_any, wgs_x, wgs_y, X0 / a.wg_w, Y0 / a.wg_h, nwgx, nwgy,
                        (uint32_t) (a.wg_w * a.wg_h));
}

#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)

void plb_fill_tail(plb_tail_args *t, const struct plb_pixel_launch *p)
{
    memset(t, 0, sizeof(*t));
    t->nplanes = p->nplanes;
    for (int i = 0; i < p->nplanes; i++) {
        t->planes[i] = p->planes[i];
        // deband/grain shaders are always evaluated on exact texel centres of their own
        // plane (1:1 sub-pass); plain sampling follows the requested coordinate mode
        t->plane_cc[i] = plb_coord_init(p->coord_mode, p->planes[i].tex, p->planes[i].rect,
                                        p->out_w, p->out_h);
    }
    memcpy(t->init, p->init, sizeof(t->init));
    for (int k = 0; k < 4; k++)
        t->src_map[k] = -1;
    t->dst = p->dst;
    t->dst_x0 = p->dst_rect.x0;
    t->dst_y0 = p->dst_rect.y0;
    t->out_w = p->out_w;
    t->out_h = p->out_h;
    t->frame_index = p->frame_index;
    t->wg_w = PLB_DEF(p->wg_w, 16);
    t->wg_h = PLB_DEF(p->wg_h, 16);
    t->prog = p->prog;
    if (p->has_dovi)
        t->dovi = p->dovi;
}

static bool prog_has_peak(const plb_color_prog *prog)
{
    for (int i = 0; i < prog->n; i++) {
        if (prog->ops[i].kind == PLB_OP_PEAK)
            return true;
    }
    return false;
}

// Light passes (plane merges, FBO copies, 1:1 or bilinear resizes with at most scale / matrix /
// swizzle steps) are pure data movement: no transcendental tables, every fetch of the thread's
// four rows is issued before the first use (memory-level parallelism instead of four dependent
// round trips), register-resident colour.
__global__ void __launch_bounds__(256)
k_pixel_light(const __grid_constant__ plb_tail_args ap)
{
    plb_tail_stage(ap);
    __syncthreads();
    const plb_tail_args &a = plb_s_tail;
    const int gx = blockIdx.x * PIX_TILE + threadIdx.x;
    const int gy0 = blockIdx.y * PIX_TILE + threadIdx.y;
    if (gx >= a.out_w)
        return;
    float color[4][4];
    #pragma unroll
    for (int r = 0; r < 4; r++) {
        #pragma unroll
        for (int k = 0; k < 4; k++)
            color[r][k] = a.init[k];
    }
    #pragma unroll 1
    for (int i = 0; i < a.nplanes; i++) {
        const plb_plane_src &pl = a.planes[i];
        float4 c4[4];
        #pragma unroll
        for (int r = 0; r < 4; r++) {
            const int gy = min(gy0 + 8 * r, a.out_h - 1);
            float posx, posy;
            plb_coord_pos(a.plane_cc[i], gx, gy, posx, posy);
            c4[r] = pl.linear ? plb_tex_linear(pl.tex, posx, posy) : plb_tex_nearest(pl.tex, posx, posy);
        }
        #pragma unroll
        for (int r = 0; r < 4; r++) {
            const float tmp[4] = { c4[r].x * pl.scale, c4[r].y * pl.scale, c4[r].z * pl.scale, c4[r].w * pl.scale };
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int m = pl.map[k];
                #pragma unroll
                for (int j = 0; j < 4; j++)
                    color[r][j] = m == j ? tmp[k] : color[r][j];
            }
        }
    }
    #pragma unroll
    for (int r = 0; r < 4; r++) {
        const int gy = gy0 + 8 * r;
        if (gy >= a.out_h)
            continue;
        plb_run_light_ops(color[r]);
        plb_store(a.dst, a.dst_x0 + gx, a.dst_y0 + gy,
                  make_float4(color[r][0], color[r][1], color[r][2], color[r][3]));
    }
}

static bool pixel_is_light(const struct plb_pixel_launch *p)
{
    if (!p->dst.ptr)
        return false;
    for (int i = 0; i < p->nplanes; i++) {
        if (p->planes[i].kind != PLB_PLANE_DIRECT)
            return false;
    }
    for (int i = 0; i < p->prog.n; i++) {
        const int k = p->prog.ops[i].kind;
        if (k != PLB_OP_SCALE && k != PLB_OP_MATRIX && k != PLB_OP_ROUND_FBO && k != PLB_OP_ALPHA &&
            k != PLB_OP_SWIZZLE)
            return false;
    }
    return true;
}

// The HDR -> SDR output pass (linearize, tone / gamut mapping, delinearize, dither + light steps
// on directly sampled planes): plb_run_hdr_ops with both transfer functions fixed at compile
// time, everything but the transcendental functions inline, colour in registers.
template <int LIN_TRC, int DELIN_TRC, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_pixel_hdr(const __grid_constant__ plb_tail_args ap)
{
    plm_tables_init();
    plb_tail_stage(ap);
    __syncthreads();
    const plb_tail_args &a = plb_s_tail;
    const int gx = blockIdx.x * PIX_TILE + threadIdx.x;
    if (gx >= a.out_w)
        return;
    #pragma unroll 1
    for (int r = 0; r < PIX_TILE / 8; r++) {
        const int gy = blockIdx.y * PIX_TILE + threadIdx.y + 8 * r;
        if (gy >= a.out_h)
            break;
        float color[4] = { a.init[0], a.init[1], a.init[2], a.init[3] };
        #pragma unroll 1
        for (int i = 0; i < a.nplanes; i++) {
            const plb_plane_src &pl = a.planes[i];
            float4 c4;
            if (a.plane_cc[i].ident && !pl.linear && plb_tex_fast_kind(pl.tex) == 3) {
                // the usual source: the rgba16hf image FBO, 1:1 (texel = pixel, always in range)
                c4 = plb_load_rgba16hf(pl.tex, gx, gy);
            } else {
                float posx, posy;
                plb_coord_pos(a.plane_cc[i], gx, gy, posx, posy);
                c4 = pl.linear ? plb_tex_linear(pl.tex, posx, posy) : plb_tex_nearest(pl.tex, posx, posy);
            }
            const float tmp[4] = { c4.x * pl.scale, c4.y * pl.scale, c4.z * pl.scale, c4.w * pl.scale };
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const int m = pl.map[k];
                #pragma unroll
                for (int j = 0; j < 4; j++)
                    color[j] = m == j ? tmp[k] : color[j];
            }
        }
        plb_run_hdr_ops<LIN_TRC, DELIN_TRC>(gx, gy, color);
        plb_store(a.dst, a.dst_x0 + gx, a.dst_y0 + gy, make_float4(color[0], color[1], color[2], color[3]));
    }
}

// Plane passes without a heavy colour program: debanding and AV1 grain synthesis into an FBO
// (renderer.c:1276-1551), with at most light steps behind. Sampling, debanding and grain are
// inline (one pixel at a time, x fastest); only sin/cos is a call.
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_pixel_planes(const __grid_constant__ plb_tail_args ap)
{
    plb_tail_stage(ap);
    __syncthreads();
    const plb_tail_args &a = plb_s_tail;
    const int gx = blockIdx.x * PIX_TILE + threadIdx.x;
    if (gx >= a.out_w)
        return;
    #pragma unroll 1
    for (int r = 0; r < PIX_TILE / 8; r++) {
        const int gy = blockIdx.y * PIX_TILE + threadIdx.y + 8 * r;
        if (gy >= a.out_h)
            break;
        float color[4] = { a.init[0], a.init[1], a.init[2], a.init[3] };
        #pragma unroll 1
        for (int i = 0; i < a.nplanes; i++) {
            const plb_plane_src &pl = a.planes[i];
            float tmp[4];
            plb_sample_plane_simple(pl, a.plane_cc[i], gx, gy, a.frame_index, tmp);
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const float v = tmp[k] * pl.scale;
                const int m = pl.map[k];
                #pragma unroll
                for (int j = 0; j < 4; j++)
                    color[j] = m == j ? v : color[j];
            }
        }
        plb_run_light_ops(color);
        plb_store(a.dst, a.dst_x0 + gx, a.dst_y0 + gy, make_float4(color[0], color[1], color[2], color[3]));
    }
}

static bool pixel_is_planes(const struct plb_pixel_launch *p)
{
    if (!p->dst.ptr)
        return false;
    for (int i = 0; i < p->nplanes; i++) {
        const int k = p->planes[i].kind;
        if (k != PLB_PLANE_DIRECT && k != PLB_PLANE_DEBAND && k != PLB_PLANE_GRAIN_AV1)
            return false;
    }
    for (int i = 0; i < p->prog.n; i++) {
        const int k = p->prog.ops[i].kind;
        if (k != PLB_OP_SCALE && k != PLB_OP_MATRIX && k != PLB_OP_ROUND_FBO && k != PLB_OP_ALPHA &&
            k != PLB_OP_SWIZZLE)
            return false;
    }
    return true; // all-direct launches were taken by k_pixel_light before
}

// Eligibility of k_pixel_hdr: direct planes only, a tone / gamut mapping step, PQ in and BT.1886
// out wherever the program (de)linearizes, nothing that needs the general interpreter.
static bool pixel_is_hdr(const struct plb_pixel_launch *p)
{
    if (!p->dst.ptr)
        return false;
    for (int i = 0; i < p->nplanes; i++) {
        if (p->planes[i].kind != PLB_PLANE_DIRECT)
            return false;
    }
    bool cmap = false;
    for (int i = 0; i < p->prog.n; i++) {
        const struct plb_color_op *op = &p->prog.ops[i];
        switch (op->kind) {
        case PLB_OP_SCALE: case PLB_OP_MATRIX: case PLB_OP_ROUND_FBO: case PLB_OP_ALPHA:
        case PLB_OP_SWIZZLE: case PLB_OP_DITHER:
            break;
        case PLB_OP_COLOR_MAP:
            cmap = true;
            break;
        case PLB_OP_LINEARIZE:
            if (op->u.trc.trc != TRC_PQ)
                return false;
            break;
        case PLB_OP_DELINEARIZE:
            if (op->u.trc.trc != TRC_BT_1886)
                return false;
            break;
        default:
            return false;
        }
    }
    return cmap;
}

extern "C" bool plb_launch_pixel(pl_gpu gpu, const struct plb_pixel_launch *p)
{
    static plb_tail_args t;
    plb_fill_tail(&t, p);
    if (pixel_is_light(p)) {
        dim3 block(32, 8);
        dim3 grid((p->out_w + PIX_TILE - 1) / PIX_TILE, (p->out_h + PIX_TILE - 1) / PIX_TILE);
        k_pixel_light<<<grid, block, 0, STREAM(gpu)>>>(t);
        plb_gpu_priv(gpu)->launches++;
        return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel_light");
    }
    static int planes_mode = -1; // PLB_PIXEL_PLANES=0 keeps the general kernel (A/B tests)
    if (planes_mode < 0) {
        const char *e = getenv("PLB_PIXEL_PLANES");
        planes_mode = e ? atoi(e) : 4;
    }
    if (planes_mode > 0 && pixel_is_planes(p)) {
        dim3 block(32, 8);
        dim3 grid((p->out_w + PIX_TILE - 1) / PIX_TILE, (p->out_h + PIX_TILE - 1) / PIX_TILE);
        // measured on B200 (chroma deband of a 4K frame, µs): 4 CTAs/SM 62, 5: 64 (spills); general kernel 107
        if (planes_mode == 3)
            k_pixel_planes<3><<<grid, block, 0, STREAM(gpu)>>>(t);
        else if (planes_mode == 5)
            k_pixel_planes<5><<<grid, block, 0, STREAM(gpu)>>>(t);
        else
            k_pixel_planes<4><<<grid, block, 0, STREAM(gpu)>>>(t);
        plb_gpu_priv(gpu)->launches++;
        return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel_planes");
    }
    static int hdr_mode = -1; // PLB_PIXEL_HDR=0 keeps the general kernel (measurements, A/B tests)
    if (hdr_mode < 0) {
        const char *e = getenv("PLB_PIXEL_HDR");
        hdr_mode = e ? atoi(e) : 5;
    }
    if (hdr_mode > 0 && pixel_is_hdr(p)) {
        dim3 block(32, 8);
        dim3 grid((p->out_w + PIX_TILE - 1) / PIX_TILE, (p->out_h + PIX_TILE - 1) / PIX_TILE);
        // measured on B200 (4K colour pass, µs): 3 CTAs/SM 923, 4: 930, 5 (48 registers): 909;
        // the general kernel: 1260
        if (hdr_mode == 3)
            k_pixel_hdr<TRC_PQ, TRC_BT_1886, 3><<<grid, block, 0, STREAM(gpu)>>>(t);
        else if (hdr_mode == 4)
            k_pixel_hdr<TRC_PQ, TRC_BT_1886, 4><<<grid, block, 0, STREAM(gpu)>>>(t);
        else
            k_pixel_hdr<TRC_PQ, TRC_BT_1886, 5><<<grid, block, 0, STREAM(gpu)>>>(t);
        plb_gpu_priv(gpu)->launches++;
        return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel_hdr");
    }
    if (prog_has_peak(&p->prog)) {
        if (PIX_TILE % t.wg_w || PIX_TILE % t.wg_h ||
            (PIX_TILE / t.wg_w) * (PIX_TILE / t.wg_h) > PLB_MAX_WG_PER_CTA) {
            PLB_ERR(gpu->log, "peak detection: unsupported work group size %dx%d", t.wg_w, t.wg_h);
            return false;
        }
    }
    dim3 block(32, 8);
    dim3 grid((p->out_w + PIX_TILE - 1) / PIX_TILE, (p->out_h + PIX_TILE - 1) / PIX_TILE);
    if (prog_has_peak(&p->prog)) {
        // cover every work group of the reference dispatch, including overhanging ones
        grid.x = ((p->out_w + t.wg_w - 1) / t.wg_w * t.wg_w + PIX_TILE - 1) / PIX_TILE;
        grid.y = ((p->out_h + t.wg_h - 1) / t.wg_h * t.wg_h + PIX_TILE - 1) / PIX_TILE;
    }
    static int minb;
    if (!minb) {
        const char *e = getenv("PLB_PIXEL_MINB");
        const int v = e ? atoi(e) : 4;
        minb = v >= 1 && v <= 5 ? v : 4;
    }
    if (minb == 1)
        k_pixel<1><<<grid, block, 0, STREAM(gpu)>>>(t);
    else if (minb == 5)
        k_pixel<5><<<grid, block, 0, STREAM(gpu)>>>(t);
    else if (minb == 4)
        k_pixel<4><<<grid, block, 0, STREAM(gpu)>>>(t);
    else if (minb == 2)
        k_pixel<2><<<grid, block, 0, STREAM(gpu)>>>(t);
    else
        k_pixel<3><<<grid, block, 0, STREAM(gpu)>>>(t);
    plb_gpu_priv(gpu)->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel");
}

// ---------------------------------------------------------------------------------------
// pl_cuda_math_eval: diagnostics front-end to plb_math.cuh
// ---------------------------------------------------------------------------------------
__global__ void k_math_eval(int fn, const float *a, const float *b, float *out, size_t n)
{
    plm_tables_init();
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n;
         i += (size_t) gridDim.x * blockDim.x) {
        float s, c, r;
        switch (fn) {
        case 0: r = plm_powf(a[i], b[i]); break;
        case 1: r = plm_expf(a[i]); break;
        case 2: r = plm_logf(a[i]); break;
        case 3: plm_sincosf(a[i], &s, &c); r = s; break;
        case 4: plm_sincosf(a[i], &s, &c); r = c; break;
        case 5: r = plm_atan2f(a[i], b[i]); break;
        case 6: plm_sincos_fast(a[i], &s, &c); r = s; break;
        case 7: plm_sincos_fast(a[i], &s, &c); r = c; break;
        case 8: r = plm_atan2_fast(a[i], b[i]); break;
        case 16 + PLB_CV_PQ_EOTF_SDR:   r = plb_pq_eotf_cv<PLB_CV_PQ_EOTF_SDR>(a[i]); break;
        case 16 + PLB_CV_PQ_EOTF:       r = plb_pq_eotf_cv<PLB_CV_PQ_EOTF>(a[i]); break;
        case 16 + PLB_CV_PQ_OETF_SDR:   r = plb_pq_oetf_cv<PLB_CV_PQ_OETF_SDR>(a[i], false); break;
        case 16 + PLB_CV_PQ_OETF_DELIN: r = plb_pq_oetf_cv<PLB_CV_PQ_OETF_DELIN>(a[i], false); break;
        case 16 + PLB_CV_PQ_OETF:       r = plb_pq_oetf_cv<PLB_CV_PQ_OETF>(a[i], false); break;
        case 16 + PLB_CV_PQX_OETF: {
            const float PQ_M1 = 2610.f / 4096 * 1.f / 4, PQ_M2 = 2523.f / 4096 * 128,
                        PQ_C1 = 3424.f / 4096, PQ_C2 = 2413.f / 4096 * 32, PQ_C3 = 2392.f / 4096 * 32;
            if (!plm_curve<PLB_CV_PQX_OETF>(a[i], r)) {
                r = plm_powf(a[i], PQ_M1);
                r = (PQ_C1 + PQ_C2 * r) / (1.0f + PQ_C3 * r);
                r = plm_powf(r, PQ_M2);
            }
            break;
        }
        case 16 + PLB_CV_POW_24:        r = plm_pow_cv<PLB_CV_POW_24>(a[i]); break;
        case 16 + PLB_CV_POW_INV24:     r = plm_pow_cv<PLB_CV_POW_INV24>(a[i]); break;
        case 16 + PLB_CV_POW_22:        r = plm_pow_cv<PLB_CV_POW_22>(a[i]); break;
        case 16 + PLB_CV_POW_INV22:     r = plm_pow_cv<PLB_CV_POW_INV22>(a[i]); break;
        default: r = __int_as_float(0x7fc00000);
        }
        out[i] = r;
    }
}

extern "C" PL_API bool pl_cuda_math_eval(pl_gpu gpu, int fn, const float *a, const float *b,
                                         float *out, size_t n)
{
    if (!n)
        return true;
    cudaStream_t st = (cudaStream_t) pl_cuda_stream(gpu);
    float *d = NULL;
    if (cudaMalloc(&d, 3 * n * sizeof(float)) != cudaSuccess)
        return false;
    bool ok = cudaMemcpyAsync(d, a, n * sizeof(float), cudaMemcpyHostToDevice, st) == cudaSuccess &&
              cudaMemcpyAsync(d + n, b ? b : a, n * sizeof(float), cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (ok) {
        k_math_eval<<<148 * 8, 256, 0, st>>>(fn, d, d + n, d + 2 * n, n);
        ok = cudaGetLastError() == cudaSuccess &&
             cudaMemcpyAsync(out, d + 2 * n, n * sizeof(float), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
    }
    cudaFree(d);
    return ok;
}
