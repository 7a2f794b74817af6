// pl_shader_deinterlace (src/shaders/deinterlacing.c:32-370): weave / bob / yadif / bwdif on
// interleaved (MBAFF) frames. One thread per output pixel; lines of the field being shown are
// copied, the others are predicted from the neighbouring lines of the current frame and, for the
// temporal algorithms, of the previous / next frames. Texels are fetched with NEAREST filtering
// and MIRROR addressing at pos + pt * (X, Y), i.e. texel (x + X, y + Y) reflected at the borders.
// Arithmetic: the GLSL statement by statement, fp32, no contraction.
#include "pixel_core.cuh"

struct deint_args {
    plb_tex_desc cur, prev, next;   // prev / next = cur when absent
    int field;                      // 0 none, 1 top (even lines kept), 2 bottom
    int prev2_is_prev, next2_is_cur;
    int algo, skip_spatial_check, intra_only;
    uint8_t comp_mask;
    plb_tail_args tail;
};

PLB_DEV float4 deint_get(const plb_tex_desc &t, int x, int y)
{
    return plb_load(t, plb_wrap_index(x, t.w, 2), plb_wrap_index(y, t.h, 2));
}

// yadif spatial predictor (:147-171); a..g = line above (x-3 .. x+3), h..n = line below
PLB_DEV float yadif_spatial(float a, float b, float c, float d, float e, float f, float g,
                            float h, float i, float j, float k, float l, float m, float n)
{
    float spatial_pred = (d + k) / 2.0f;
    float spatial_score = fabsf(c - j) + fabsf(d - k) + fabsf(e - l) - 0.003922f; // "%f" of 1/255
    float score = fabsf(b - k) + fabsf(c - l) + fabsf(d - m);
    if (score < spatial_score) {
        spatial_pred = (c + l) / 2.0f;
        spatial_score = score;
        score = fabsf(a - l) + fabsf(b - m) + fabsf(c - n);
        if (score < spatial_score) {
            spatial_pred = (b + m) / 2.0f;
            spatial_score = score;
        }
    }
    score = fabsf(d - i) + fabsf(e - j) + fabsf(f - k);
    if (score < spatial_score) {
        spatial_pred = (e + j) / 2.0f;
        spatial_score = score;
        score = fabsf(e - h) + fabsf(f - i) + fabsf(g - j);
        if (score < spatial_score) {
            spatial_pred = (f + i) / 2.0f;
            spatial_score = score;
        }
    }
    return spatial_pred;
}

// yadif temporal predictor (:205-230)
PLB_DEV float yadif_temporal(float A, float B, float C, float D, float E, float F, float G, float H,
                             float I, float J, float K, float L, float spatial_pred, bool skip_check)
{
    const float p0 = (C + H) / 2.0f, p1 = F, p2 = (D + I) / 2.0f, p3 = G, p4 = (E + J) / 2.0f;
    const float tdiff0 = fabsf(D - I) / 2.0f;
    const float tdiff1 = (fabsf(A - F) + fabsf(B - G)) / 2.0f;
    const float tdiff2 = (fabsf(K - F) + fabsf(G - L)) / 2.0f;
    float diff = fmaxf(tdiff0, fmaxf(tdiff1, tdiff2));
    if (!skip_check) {
        const float maxi = fmaxf(p2 - fminf(p3, p1), fminf(p0 - p1, p4 - p3));
        const float mini = fminf(p2 - fmaxf(p3, p1), fmaxf(p0 - p1, p4 - p3));
        diff = fmaxf(diff, fmaxf(mini, -maxi));
    }
    if (spatial_pred > p2 + diff)
        spatial_pred = p2 + diff;
    if (spatial_pred < p2 - diff)
        spatial_pred = p2 - diff;
    return spatial_pred;
}

// bwdif (:262-303); cur = {-3, -1, +1, +3}, prev / next = {-1, +1}, prev2 / next2 = {-4 .. +4}
PLB_DEV float bwdif_process(const float cur[4], const float prev[2], const float next[2],
                            const float prev2[5], const float next2[5])
{
    const float lf0 = 4309.0f / 8192.0f, lf1 = 213.0f / 8192.0f;
    const float hf0 = 5570.0f / 8192.0f, hf1 = 3801.0f / 8192.0f, hf2 = 1016.0f / 8192.0f;
    const float sp0 = 5077.0f / 8192.0f, sp1 = 981.0f / 8192.0f;
    const float s = prev2[2] + next2[2];
    const float d = s / 2.0f;
    const float c = cur[1], e = cur[2];
    const float tdiff0 = fabsf(prev2[2] - next2[2]);
    const float tdiff1 = fabsf(prev[0] - c) + fabsf(prev[1] - e);
    const float tdiff2 = fabsf(next[0] - c) + fabsf(next[1] - e);
    float diff = fmaxf(tdiff0, fmaxf(tdiff1, tdiff2)) / 2.0f;
    const bool diff_mask = diff == 0.0f;
    const float bs = prev2[1] + next2[1], fs = prev2[3] + next2[3];
    const float b = (bs / 2.0f) - c, f = (fs / 2.0f) - c;
    const float dc = d - c, de = d - e;
    const float mmax = fmaxf(de, fmaxf(dc, fminf(b, f)));
    const float mmin = fminf(de, fminf(dc, fmaxf(b, f)));
    diff = fmaxf(diff, fmaxf(mmin, -mmax));
    const float single = sp0 * (c + e) - sp1 * (cur[0] + cur[3]);
    float all = (hf0 * s - hf1 * (bs + fs) + hf2 * (prev2[0] + next2[0] + prev2[4] + next2[4])) / 4.0f;
    all += lf0 * (c + e) - lf1 * (cur[0] + cur[3]);
    const bool mask = fabsf(c - e) > tdiff0;
    float interpol = mask ? all : single;
    interpol = fminf(fmaxf(interpol, d - diff), d + diff);
    return diff_mask ? d : interpol;
}

PLB_DEV float bwdif_intra(const float cur[4])
{
    const float sp0 = 5077.0f / 8192.0f, sp1 = 981.0f / 8192.0f;
    return sp0 * (cur[1] + cur[2]) - sp1 * (cur[0] + cur[3]);
}

PLB_DEV float comp_of(const float4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

PLB_DEV void deint_pixel(const deint_args &a, int x, int y, float res[4])
{
    const float4 c0 = deint_get(a.cur, x, y);
    res[0] = c0.x; res[1] = c0.y; res[2] = c0.z; res[3] = c0.w;
    if (a.field == 0 || (y % 2) == (a.field == 1 ? 0 : 1) || a.algo == 0 /* weave */)
        return;
    if (a.algo == 1) { // bob
        const float4 v = deint_get(a.cur, x, y + (a.field == 1 ? -1 : 1));
        res[0] = v.x; res[1] = v.y; res[2] = v.z; res[3] = v.w;
        return;
    }
    const plb_tex_desc &prev2 = a.prev2_is_prev ? a.prev : a.cur;
    const plb_tex_desc &next2 = a.next2_is_cur ? a.cur : a.next;
    if (a.algo == 2) { // yadif
        float4 up[7], dn[7];
        #pragma unroll
        for (int i = 0; i < 7; i++) {
            up[i] = deint_get(a.cur, x + i - 3, y - 1);
            dn[i] = deint_get(a.cur, x + i - 3, y + 1);
        }
        const float4 A = deint_get(a.prev, x, y - 1), B = deint_get(a.prev, x, y + 1);
        const float4 C = deint_get(prev2, x, y - 2), D = deint_get(prev2, x, y), E = deint_get(prev2, x, y + 2);
        const float4 F = up[3], G = dn[3];
        const float4 H = deint_get(next2, x, y - 2), I = deint_get(next2, x, y), J = deint_get(next2, x, y + 2);
        const float4 K = deint_get(a.next, x, y - 1), L = deint_get(a.next, x, y + 1);
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const float sp = yadif_spatial(comp_of(up[0], k), comp_of(up[1], k), comp_of(up[2], k), comp_of(up[3], k),
                                           comp_of(up[4], k), comp_of(up[5], k), comp_of(up[6], k),
                                           comp_of(dn[0], k), comp_of(dn[1], k), comp_of(dn[2], k), comp_of(dn[3], k),
                                           comp_of(dn[4], k), comp_of(dn[5], k), comp_of(dn[6], k));
            res[k] = yadif_temporal(comp_of(A, k), comp_of(B, k), comp_of(C, k), comp_of(D, k), comp_of(E, k),
                                    comp_of(F, k), comp_of(G, k), comp_of(H, k), comp_of(I, k), comp_of(J, k),
                                    comp_of(K, k), comp_of(L, k), sp, a.skip_spatial_check != 0);
        }
        return;
    }
    // bwdif
    const float4 cu[4] = { deint_get(a.cur, x, y - 3), deint_get(a.cur, x, y - 1),
                           deint_get(a.cur, x, y + 1), deint_get(a.cur, x, y + 3) };
    if (a.intra_only) {
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const float c4[4] = { comp_of(cu[0], k), comp_of(cu[1], k), comp_of(cu[2], k), comp_of(cu[3], k) };
            res[k] = bwdif_intra(c4);
        }
        return;
    }
    const float4 pv[2] = { deint_get(a.prev, x, y - 1), deint_get(a.prev, x, y + 1) };
    const float4 nx[2] = { deint_get(a.next, x, y - 1), deint_get(a.next, x, y + 1) };
    float4 p2[5], n2[5];
    #pragma unroll
    for (int i = 0; i < 5; i++) {
        p2[i] = deint_get(prev2, x, y + 2 * i - 4);
        n2[i] = deint_get(next2, x, y + 2 * i - 4);
    }
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        const float c4[4] = { comp_of(cu[0], k), comp_of(cu[1], k), comp_of(cu[2], k), comp_of(cu[3], k) };
        const float pr[2] = { comp_of(pv[0], k), comp_of(pv[1], k) }, ne[2] = { comp_of(nx[0], k), comp_of(nx[1], k) };
        float q2[5], m2[5];
        #pragma unroll
        for (int i = 0; i < 5; i++) {
            q2[i] = comp_of(p2[i], k);
            m2[i] = comp_of(n2[i], k);
        }
        res[k] = bwdif_process(c4, pr, ne, q2, m2);
    }
}

__global__ void __launch_bounds__(256)
k_deinterlace(const __grid_constant__ deint_args a)
{
    plm_tables_init();
    plb_tail_stage(a.tail);
    __syncthreads();
    const plb_tail_args &t = plb_s_tail;
    const int gx = blockIdx.x * 32 + threadIdx.x;
    if (gx >= t.out_w)
        return;
    #pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const int gy = blockIdx.y * 32 + threadIdx.y + 8 * r;
        if (gy >= t.out_h)
            break;
        float res[4];
        deint_pixel(a, gx, gy, res);
        // vec4 color = vec4(0, 0, 0, 1); color.<mask> = res
        float prim[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (a.comp_mask & (1 << k))
                prim[k] = res[k];
        }
        float color[4];
        uint32_t ypq = 0;
        const plb_op_peak *pk = nullptr;
        plb_tail_color(gx, gy, prim, color, &ypq, &pk);
        if (t.dst.ptr)
            plb_store(t.dst, t.dst_x0 + gx, t.dst_y0 + gy, make_float4(color[0], color[1], color[2], color[3]));
    }
}

#define STREAM(gpu) ((cudaStream_t) plb_gpu_priv(gpu)->stream)
void plb_fill_tail(plb_tail_args *t, const struct plb_pixel_launch *p);
PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_deint)

extern "C" bool plb_launch_deint(pl_gpu gpu, const struct plb_deint_launch *p)
{
    const struct plb_pixel_launch *tl = &p->tail;
    for (int i = 0; i < tl->prog.n; i++) {
        if (tl->prog.ops[i].kind == PLB_OP_PEAK) {
            PLB_ERR(gpu->log, "deinterlacing: peak detection cannot be merged into this pass");
            return false;
        }
    }
    deint_args a;
    memset(&a, 0, sizeof(a));
    a.cur = p->cur;
    a.prev = p->prev;
    a.next = p->next;
    a.field = p->field;
    a.prev2_is_prev = p->prev2_is_prev;
    a.next2_is_cur = p->next2_is_cur;
    a.algo = p->algo;
    a.skip_spatial_check = p->skip_spatial_check;
    a.intra_only = p->intra_only;
    a.comp_mask = p->comp_mask;
    plb_fill_tail(&a.tail, tl);
    memcpy(a.tail.src_map, p->src_map, sizeof(p->src_map));
    dim3 block(32, 8);
    dim3 grid((tl->out_w + 31) / 32, (tl->out_h + 31) / 32);
    k_deinterlace<<<grid, block, 0, STREAM(gpu)>>>(a);
    plb_gpu_priv(gpu)->launches++;
    return plb_cuda_check(gpu, cudaGetLastError(), "k_deinterlace");
}
