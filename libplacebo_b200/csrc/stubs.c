// Entry points that are declared in the public header but not implemented yet fail loudly.
#include "shader_priv.h"

void plb_grain_obj_free(pl_gpu gpu, struct plb_grain_obj *g) { (void) gpu; (void) g; }

bool pl_needs_film_grain(const struct pl_film_grain_params *params) { (void) params; return false; }

bool pl_shader_film_grain(pl_shader sh, pl_shader_obj *grain_state,
                          const struct pl_film_grain_params *params)
{
    (void) grain_state; (void) params;
    plb_shader_fail(sh, "film grain: not implemented yet");
    return false;
}

bool plb_record_swizzle(pl_shader sh, int comps, const int comp_map[4], bool force_alpha)
{
    static const int def_map[4] = {0, 1, 2, 3};
    comp_map = comp_map ? comp_map : def_map;
    struct plb_color_op *op = plb_shader_push_op(sh, PLB_OP_SWIZZLE);
    if (!op)
        return false;
    for (int k = 0; k < 4; k++)
        op->u.swizzle.map[k] = comp_map[k];
    op->u.swizzle.comps = comps;
    op->u.swizzle.force_alpha = force_alpha;
    return true;
}

bool plb_plane_film_grain(pl_renderer rr, pl_shader sh, pl_shader_obj *state,
                          const struct pl_frame *image, const struct pl_plane *plane,
                          pl_tex tex, pl_tex luma_tex, int luma_comp, struct pl_color_repr *repr)
{
    (void) rr; (void) state; (void) image; (void) plane; (void) tex; (void) luma_tex;
    (void) luma_comp; (void) repr;
    plb_shader_fail(sh, "film grain: not implemented yet");
    return false;
}

