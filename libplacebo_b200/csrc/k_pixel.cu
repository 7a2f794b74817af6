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
    for (int i = 0; i < t->prog.n; i++) { // the feature-map position is relative to the output
        if (t->prog.ops[i].kind == PLB_OP_COLOR_MAP) {
            t->prog.ops[i].u.cmap.out_w = p->out_w;
            t->prog.ops[i].u.cmap.out_h = p->out_h;
        }
    }
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
            if (a.plane_cc[i].ident) { // sample positions are the texel centres
                c4[r] = plb_load_clamped(pl.tex, gx, gy);
                continue;
            }
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

// The HDR -> SDR output pass (pass_convert_colors + pass_output_target, renderer.c:2157-2279,
// 2586-2964): the rgba16hf image FBO 1:1 -> [light steps] -> linearize (PQ) -> tone / gamut mapping
// -> delinearize (BT.1886) -> [light steps] -> [dither] -> [light steps] -> target.
// Straight-line code: no op loop around the four heavy steps, the launch description is read from
// the kernel parameters (constant bank operands, no shared-memory staging), transfer functions are
// curve evaluations (plb_math.cuh). Per-pixel arithmetic is that of plb_run_color_ops, step for
// step, so the general kernel (PLB_PIXEL_HDR=0) produces the same bytes.
#define HDR_MAX_LIGHT 6
struct plb_hdr_args {
    plb_tex_desc src, dst;
    int dst_x0, dst_y0, out_w, out_h;
    float src_scale;
    // light steps between delinearize and dither / after dither that are not folded below
    int n_mid, n_post;
    int has_dither;
    // folded tail: [scale] -> swizzle (identity selection: absent components take 0, 0, 0, 1)
    int fold_post;                     // n_post == 0 and the fields below describe the tail
    float post_scale; int post_scale_alpha, has_post_scale;
    int swz_keep, has_swz, swz_force_alpha;
    plb_trc_consts lin, delin;
    plb_op_color_map cmap;
    plb_op_dither dither;
    plb_color_op light[HDR_MAX_LIGHT];
};

PLB_DEV float4 hdr_unpack(uint2 raw)
{
    const float2 a = __half22float2(*(__half2 *) &raw.x), b = __half22float2(*(__half2 *) &raw.y);
    return make_float4(a.x, a.y, b.x, b.y);
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_pixel_hdr(const __grid_constant__ plb_hdr_args a)
{
    plm_tables_init(); // operands of the exact (out-of-domain) paths
    __syncthreads();
    const int gx = blockIdx.x * PIX_TILE + threadIdx.x;
    const int gy0 = blockIdx.y * PIX_TILE + threadIdx.y;
    if (gx >= a.out_w || gy0 >= a.out_h)
        return;
    const char *src = (const char *) a.src.ptr + (size_t) gx * 8;
    uint2 raw = *(const uint2 *) (src + (size_t) gy0 * a.src.pitch);
    #pragma unroll 1
    for (int r = 0; r < PIX_TILE / 8; r++) {
        const int gy = gy0 + 8 * r;
        if (gy >= a.out_h)
            break;
        const float4 t = hdr_unpack(raw);
        if (gy + 8 < a.out_h && r + 1 < PIX_TILE / 8) // next row's texel while this one computes
            raw = *(const uint2 *) (src + (size_t) (gy + 8) * a.src.pitch);
        float c[4] = { t.x * a.src_scale, t.y * a.src_scale, t.z * a.src_scale, t.w * a.src_scale };

        // pl_shader_linearize, PQ (colorspace.c:662-668): clamp, EOTF scaled to 203 nits = 1
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = fmaxf(c[k], 0.0f);
        plb_curve3_s<PLB_CV_PQ_EOTF_SDR>(c);
        plb_color_map<true>(a.cmap, c, gx, gy);
        // pl_shader_delinearize, BT.1886 (colorspace.c:759-767)
        if (a.delin.black_scaled) {
            const float ps_a = 1 / (a.delin.csp_max - a.delin.csp_min),
                        ps_b = -a.delin.csp_min / (a.delin.csp_max - a.delin.csp_min);
            #pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = ps_a * c[k] + ps_b;
        }
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = a.delin.a * fmaxf(c[k], 0.0f);
        plb_curve3_s<PLB_CV_POW_INV24>(c);
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] -= a.delin.b;

        if (a.has_dither)
            plb_dither(a.dither, gx, gy, c);
        {
            if (a.has_post_scale) {
                c[0] *= a.post_scale; c[1] *= a.post_scale; c[2] *= a.post_scale;
                if (a.post_scale_alpha)
                    c[3] *= a.post_scale;
            }
            if (a.has_swz) {
                const float o3 = c[3];
                if (!(a.swz_keep & 1)) c[0] = 0.0f;
                if (!(a.swz_keep & 2)) c[1] = 0.0f;
                if (!(a.swz_keep & 4)) c[2] = 0.0f;
                if (!(a.swz_keep & 8)) c[3] = 1.0f;
                if (a.swz_force_alpha)
                    c[3] = o3;
            }
        }
        const int ox = a.dst_x0 + gx, oy = a.dst_y0 + gy;
        if (a.dst.kind == PLB_FK_UNORM8 && a.dst.comps == 4) {
            *(uint32_t *) ((char *) a.dst.ptr + (size_t) oy * a.dst.pitch + (size_t) ox * 4) =
                plb_unorm(c[0], 255.0f) | (plb_unorm(c[1], 255.0f) << 8) |
                (plb_unorm(c[2], 255.0f) << 16) | (plb_unorm(c[3], 255.0f) << 24);
        } else if (a.dst.kind == PLB_FK_UNORM16 && a.dst.comps == 4) {
            uint2 v;
            v.x = plb_unorm(c[0], 65535.0f) | (plb_unorm(c[1], 65535.0f) << 16);
            v.y = plb_unorm(c[2], 65535.0f) | (plb_unorm(c[3], 65535.0f) << 16);
            *(uint2 *) ((char *) a.dst.ptr + (size_t) oy * a.dst.pitch + (size_t) ox * 8) = v;
        } else {
            plb_store(a.dst, ox, oy, make_float4(c[0], c[1], c[2], c[3]));
        }
    }
}

static bool op_is_light(int k)
{
    return k == PLB_OP_SCALE || k == PLB_OP_MATRIX || k == PLB_OP_ROUND_FBO || k == PLB_OP_ALPHA ||
           k == PLB_OP_SWIZZLE;
}

// Eligibility of k_pixel_hdr and its argument block: one rgba16hf plane sampled 1:1 with the
// identity component map; program = LINEARIZE(PQ) COLOR_MAP DELINEARIZE(BT.1886) light*
// [DITHER] light*. The usual tail (scale, identity swizzle) is folded into flags.
static bool pixel_hdr_args(pl_gpu gpu, const struct plb_pixel_launch *p, const plb_tail_args *t, plb_hdr_args *h)
{
    if (!p->dst.ptr || p->nplanes != 1 || plb_gpu_priv(gpu)->exact_math)
        return false;
    const plb_plane_src &pl = p->planes[0];
    if (pl.kind != PLB_PLANE_DIRECT || !t->plane_cc[0].ident || pl.tex.ptr2 || pl.tex.addr ||
        pl.tex.kind != PLB_FK_HALF || pl.tex.comps != 4)
        return false;
    for (int k = 0; k < 4; k++) {
        if (pl.map[k] != k)
            return false;
    }
    memset(h, 0, sizeof(*h));
    const plb_color_prog &pr = p->prog;
    int i = 0, nl = 0;
    if (pr.n < 3 || pr.ops[0].kind != PLB_OP_LINEARIZE || pr.ops[0].u.trc.trc != TRC_PQ ||
        pr.ops[1].kind != PLB_OP_COLOR_MAP || pr.ops[2].kind != PLB_OP_DELINEARIZE ||
        pr.ops[2].u.trc.trc != TRC_BT_1886)
        return false;
    h->lin = pr.ops[0].u.trc;
    h->cmap = t->prog.ops[1].u.cmap; // (with the output size filled in by plb_fill_tail)
    h->delin = pr.ops[2].u.trc;
    i = 3;
    for (; i < pr.n && op_is_light(pr.ops[i].kind); i++) {
        if (nl == HDR_MAX_LIGHT)
            return false;
        h->light[nl++] = pr.ops[i];
        h->n_mid++;
    }
    if (i < pr.n && pr.ops[i].kind == PLB_OP_DITHER) {
        h->dither = pr.ops[i].u.dither;
        h->has_dither = 1;
        i++;
    } else {
        h->n_post = h->n_mid; // no dither: one run of light steps
        h->n_mid = 0;
    }
    const int post0 = i;
    for (; i < pr.n && op_is_light(pr.ops[i].kind); i++) {
        if (nl == HDR_MAX_LIGHT)
            return false;
        h->light[nl++] = pr.ops[i];
        h->n_post++;
    }
    if (i != pr.n)
        return false;
    // fold "[scale] [identity swizzle]" (what pass_output_target records)
    if (h->has_dither || h->n_post == pr.n - post0) {
        int j = post0;
        plb_hdr_args f = *h;
        f.fold_post = 1;
        if (j < pr.n && pr.ops[j].kind == PLB_OP_SCALE) {
            f.has_post_scale = 1;
            f.post_scale = pr.ops[j].u.scale.s;
            f.post_scale_alpha = pr.ops[j].u.scale.with_alpha;
            j++;
        }
        bool ok = true;
        if (j < pr.n && pr.ops[j].kind == PLB_OP_SWIZZLE) {
            const plb_color_op &op = pr.ops[j];
            f.has_swz = 1;
            f.swz_force_alpha = op.u.swizzle.force_alpha;
            for (int k = 0; k < 4; k++) {
                const int m = op.u.swizzle.map[k];
                const int sel = (k < op.u.swizzle.comps && m >= 0) ? m : -1;
                if (sel == k)
                    f.swz_keep |= 1 << k;
                else if (sel >= 0)
                    ok = false; // a real permutation: generic step
            }
            j++;
        }
        if (ok && j == pr.n) {
            f.n_post = 0;
            *h = f;
        }
    }
    // anything else between delinearize and the target (encode matrices, permutations, alpha
    // conversions) takes the general kernel: this one keeps the colour in registers throughout
    if (h->n_mid || h->n_post || !h->fold_post)
        return false;
    h->src = pl.tex;
    h->dst = p->dst;
    h->dst_x0 = p->dst_rect.x0;
    h->dst_y0 = p->dst_rect.y0;
    h->out_w = p->out_w;
    h->out_h = p->out_h;
    h->src_scale = pl.scale;
    return true;
}

// Feature extraction of contrast recovery (get_feature_map, renderer.c:2089-2154): rgba16hf image
// sampled 1:1 -> [LINEARIZE(PQ)] FEATURES -> r16hf intensity plane. Same arithmetic as the
// general kernel's steps (color_ops.cuh), curves with immediates.
struct plb_feat_args {
    plb_tex_desc src, dst;
    int dst_x0, dst_y0, out_w, out_h;
    float src_scale;
    int linearize;
    float m[9];
};

__global__ void __launch_bounds__(256)
k_pixel_feat(const __grid_constant__ plb_feat_args a)
{
    plm_tables_init();
    __syncthreads();
    const int gx = blockIdx.x * PIX_TILE + threadIdx.x;
    const int gy0 = blockIdx.y * PIX_TILE + threadIdx.y;
    if (gx >= a.out_w || gy0 >= a.out_h)
        return;
    const char *src = (const char *) a.src.ptr + (size_t) gx * 8;
    #pragma unroll 2
    for (int r = 0; r < PIX_TILE / 8; r++) {
        const int gy = gy0 + 8 * r;
        if (gy >= a.out_h)
            break;
        const float4 t = hdr_unpack(*(const uint2 *) (src + (size_t) gy * a.src.pitch));
        float c[3] = { t.x * a.src_scale, t.y * a.src_scale, t.z * a.src_scale };
        if (a.linearize) {
            #pragma unroll
            for (int k = 0; k < 3; k++)
                c[k] = fmaxf(c[k], 0.0f);
            plb_curve3_s<PLB_CV_PQ_EOTF_SDR>(c);
        }
        plb_matv(a.m, c[0], c[1], c[2]);
        #pragma unroll
        for (int k = 0; k < 3; k++)
            c[k] = fmaxf(c[k], 0.0f);
        plb_curve3_s<PLB_CV_PQ_OETF>(c);
        const float I = __fmaf_rn(0.2f, c[2], __fmaf_rn(0.4f, c[1], 0.4f * c[0]));
        *(__half *) ((char *) a.dst.ptr + (size_t) (a.dst_y0 + gy) * a.dst.pitch +
                     (size_t) (a.dst_x0 + gx) * 2) = __float2half_rn(I);
    }
}

static bool pixel_feat_args(pl_gpu gpu, const struct plb_pixel_launch *p, const plb_tail_args *t, plb_feat_args *h)
{
    if (!p->dst.ptr || p->nplanes != 1 || plb_gpu_priv(gpu)->exact_math)
        return false;
    const plb_plane_src &pl = p->planes[0];
    if (pl.kind != PLB_PLANE_DIRECT || !t->plane_cc[0].ident || pl.tex.ptr2 || pl.tex.addr ||
        pl.tex.kind != PLB_FK_HALF || pl.tex.comps != 4)
        return false;
    for (int k = 0; k < 3; k++) {
        if (pl.map[k] != k)
            return false;
    }
    if (p->dst.kind != PLB_FK_HALF || p->dst.comps != 1 || p->dst.ptr2)
        return false;
    const plb_color_prog &pr = p->prog;
    memset(h, 0, sizeof(*h));
    int i = 0;
    if (pr.n >= 1 && pr.ops[0].kind == PLB_OP_LINEARIZE && pr.ops[0].u.trc.trc == TRC_PQ) {
        h->linearize = 1;
        i = 1;
    }
    if (pr.n != i + 1 || pr.ops[i].kind != PLB_OP_FEATURES)
        return false;
    memcpy(h->m, pr.ops[i].u.features.m, sizeof(h->m));
    h->src = pl.tex;
    h->dst = p->dst;
    h->dst_x0 = p->dst_rect.x0;
    h->dst_y0 = p->dst_rect.y0;
    h->out_w = p->out_w;
    h->out_h = p->out_h;
    h->src_scale = pl.scale;
    return true;
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

extern "C" bool plb_launch_pixel(pl_gpu gpu, const struct plb_pixel_launch *p)
{
    plb_tail_args t; // on the stack: two pl_gpu may launch from two threads
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
        hdr_mode = e ? atoi(e) : 1;
    }
    plb_hdr_args h;
    if (hdr_mode > 0 && pixel_hdr_args(gpu, p, &t, &h)) {
        dim3 block(32, 8);
        dim3 grid((p->out_w + PIX_TILE - 1) / PIX_TILE, (p->out_h + PIX_TILE - 1) / PIX_TILE);
        static int hdr_minb; // PLB_HDR_MINB=4|5: resident CTAs per SM the registers are capped for
        if (!hdr_minb) {
            const char *e = getenv("PLB_HDR_MINB");
            hdr_minb = e && atoi(e) == 4 ? 4 : 5;
        }
        if (hdr_minb == 4)
            k_pixel_hdr<4><<<grid, block, 0, STREAM(gpu)>>>(h);
        else
            k_pixel_hdr<5><<<grid, block, 0, STREAM(gpu)>>>(h);
        plb_gpu_priv(gpu)->launches++;
        return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel_hdr");
    }
    plb_feat_args fa;
    if (hdr_mode > 0 && pixel_feat_args(gpu, p, &t, &fa)) {
        dim3 block(32, 8);
        dim3 grid((p->out_w + PIX_TILE - 1) / PIX_TILE, (p->out_h + PIX_TILE - 1) / PIX_TILE);
        k_pixel_feat<<<grid, block, 0, STREAM(gpu)>>>(fa);
        plb_gpu_priv(gpu)->launches++;
        return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel_feat");
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

// Frame mixing (pl_render_image_mix, renderer.c:3906-3966): texel (x, y) of every cached frame,
// linearised when several frames take part, accumulated as `mix += w * color` (fma), then the
// common tail (delinearisation and the output stage recorded behind it).
struct plb_mix_args {
    int n, linearize;
    plb_tex_desc tex[PLB_MAX_MIX];
    float weight[PLB_MAX_MIX];
    plb_trc_consts lin;
    plb_tail_args tail;
};

__global__ void __launch_bounds__(256)
k_pixel_mix(const __grid_constant__ plb_mix_args a)
{
    plm_tables_init();
    plb_tail_stage(a.tail);
    __syncthreads();
    const plb_tail_args &t = plb_s_tail;
    const int gx = blockIdx.x * PIX_TILE + threadIdx.x;
    if (gx >= t.out_w)
        return;
    #pragma unroll 1
    for (int r = 0; r < PIX_TILE / 8; r++) {
        const int gy = blockIdx.y * PIX_TILE + threadIdx.y + 8 * r;
        if (gy >= t.out_h)
            break;
        float prim[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int i = 0; i < a.n; i++) {
            const float4 c4 = plb_load(a.tex[i], gx, gy);
            float c[4] = { c4.x, c4.y, c4.z, c4.w };
            if (!a.linearize) { // a single frame: nothing to mix
                prim[0] = c[0]; prim[1] = c[1]; prim[2] = c[2]; prim[3] = c[3];
                break;
            }
            plb_linearize(a.lin.trc, a.lin, c);
            #pragma unroll
            for (int k = 0; k < 4; k++)
                prim[k] = __fmaf_rn(a.weight[i], c[k], prim[k]);
        }
        float color[4];
        uint32_t ypq = 0;
        const plb_op_peak *pk = nullptr;
        plb_tail_color(gx, gy, prim, color, &ypq, &pk);
        if (t.dst.ptr)
            plb_store(t.dst, t.dst_x0 + gx, t.dst_y0 + gy, make_float4(color[0], color[1], color[2], color[3]));
    }
}

extern "C" bool plb_launch_mix(pl_gpu gpu, const struct plb_mix_launch *p)
{
    const struct plb_pixel_launch *tl = &p->tail;
    for (int i = 0; i < tl->prog.n; i++) {
        if (tl->prog.ops[i].kind == PLB_OP_PEAK) {
            PLB_ERR(gpu->log, "frame mixing: peak detection cannot follow the mixing stage");
            return false;
        }
    }
    plb_mix_args a;
    memset(&a, 0, sizeof(a));
    a.n = p->n;
    a.linearize = p->linearize;
    memcpy(a.tex, p->tex, sizeof(a.tex));
    memcpy(a.weight, p->weight, sizeof(a.weight));
    a.lin = p->lin;
    plb_fill_tail(&a.tail, tl);
    memcpy(a.tail.src_map, p->src_map, sizeof(p->src_map));
    dim3 block(32, 8);
    dim3 grid((tl->out_w + PIX_TILE - 1) / PIX_TILE, (tl->out_h + PIX_TILE - 1) / PIX_TILE);
    k_pixel_mix<<<grid, block, 0, STREAM(gpu)>>>(a);
    plb_gpu_priv(gpu)->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_pixel_mix");
}
