// Launch interface of the EWA 2x phase kernels (k_up2.cuh), which are instantiated in the
// k_up2_*.cu translation units (split by component count: nvcc compiles them in parallel).
#pragma once
#include "pixel_core.cuh"

#define UP2_TX 16           // threads per CTA in x
#define UP2_TY 16
#define UP2_OX 8            // outputs per thread in x (four source cells)
#define UP2_OY 2
#define UP2_TW (UP2_TX * UP2_OX)               // 128 output columns per CTA
#define UP2_TH (UP2_TY * UP2_OY)               // 32 output rows
#define UP2_TILE_W (UP2_TW / 2 + 8)            // 72 source columns (multiple of 4: float4 reads)
#define UP2_TILE_H (UP2_TH / 2 + 8)            // 24 source rows

struct polar_up2_args {
    plb_tex_desc src;
    float w[2][2][8][8];    // [phase_y][phase_x][oy+3][ox+3]; phase 0 = fraction 0.75
    float norm[2][2];       // scale / wsum per phase
    int use_tma;            // interior tiles come in through the tensor map
    int src_kind;           // UP2_SRC_*: source layout with a dedicated tile loader
    // folded tail of the canonical light pass (MODE 1)
    int ep_base, ep_luma, ep_fill;
    int ep_has_mat, ep_mat_has_c; float ep_mat[9], ep_mat_c[3];
    int ep_has_scale, ep_scale_alpha; float ep_scale;
    int ep_has_swz, ep_swz_keep, ep_swz_force_alpha;
    plb_tail_args tail;
};

// Sub-texel phases of an exact 2x upscale, per axis. MODE 0 ("centred": rect origin 0, e.g.
// RGB frames or centre-sited chroma): fractions 3/4, 1/4 and the first output of a cell starts
// one texel to the left. MODE 1 ("left"/"top" sited chroma, rect origin 0.25): fractions 0, 1/2,
// both outputs share the cell's texel as base.
__host__ __device__ constexpr int up2_frac4(int mode, int phase)
{
    return mode == 0 ? (phase ? 1 : 3) : (phase ? 2 : 0);
}
__host__ __device__ constexpr int up2_shift(int mode, int phase)
{
    return mode == 0 && phase == 0 ? 1 : 0;
}

// 1 = launched, 0 = launch failed, -1 = no kernel for this combination (use the generic one).
// (xm, ym): phase mode per axis, 0 = centred, 1 = left / top sited chroma.
extern "C" int plb_up2_launch_a(pl_gpu gpu, polar_up2_args *args, int nc, int xm, int ym); // 1-2 comps, centred
extern "C" int plb_up2_launch_b(pl_gpu gpu, polar_up2_args *args, int nc, int xm, int ym); // 2 comps, sited chroma
extern "C" int plb_up2_launch_c(pl_gpu gpu, polar_up2_args *args, int nc, int xm, int ym); // 3-4 comps, centred
