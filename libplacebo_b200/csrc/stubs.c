// Entry points that are declared in the public header but not implemented yet fail loudly.
#include "shader_priv.h"

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

