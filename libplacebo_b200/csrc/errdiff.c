// pl_shader_error_diffusion host side (src/shaders/dithering.c:292-400) and the renderer's
// pass_error_diffusion (src/renderer.c:2282-2343).
#include "shader_priv.h"

static int rightmost_shifted_column(const struct pl_error_diffusion_kernel *k)
{
    int ret = 0;
    for (int y = 0; y <= PL_EDF_MAX_DY; y++) {
        for (int x = PL_EDF_MIN_DX; x <= PL_EDF_MAX_DX; x++) {
            if (k->pattern[y][x - PL_EDF_MIN_DX])
                ret = PLB_MAX(ret, x + y * k->shift);
        }
    }
    return ret;
}

size_t pl_error_diffusion_shmem_req(const struct pl_error_diffusion_kernel *kernel, int height)
{
    // PL_EDF_MAX_DY empty rows at the bottom take the errors leaving the image
    const int rows = height + PL_EDF_MAX_DY;
    const int cols = rightmost_shifted_column(kernel) + 1;
    return (size_t) rows * cols * sizeof(uint32_t);
}

bool pl_shader_error_diffusion(pl_shader sh, const struct pl_error_diffusion_params *params)
{
    if (!params || !params->input_tex || !params->output_tex) {
        plb_shader_fail(sh, "pl_shader_error_diffusion: input and output textures required");
        return false;
    }
    const int width = params->input_tex->params.w, height = params->input_tex->params.h;
    const struct pl_error_diffusion_kernel *kernel =
        params->kernel ? params->kernel : &pl_error_diffusion_sierra_lite;
    if (params->output_tex->params.w != width || params->output_tex->params.h != height) {
        plb_shader_fail(sh, "pl_shader_error_diffusion: texture sizes differ");
        return false;
    }
    if (sh->failed || sh->src != PLB_SRC_NONE) {
        plb_shader_fail(sh, "pl_shader_error_diffusion: shader must be empty (sh_require "
                        "PL_SHADER_SIG_NONE)");
        return false;
    }
    if (params->new_depth <= 0 || params->new_depth > 256) {
        PLB_WARN(sh->log, "Invalid dither depth: %d.. ignoring", params->new_depth);
        return false;
    }
    if (params->new_depth > 24) {
        plb_shader_fail(sh, "error diffusion to %d bits is not meaningful in fp32", params->new_depth);
        return false;
    }
    if (!params->output_tex->params.storable) {
        plb_shader_fail(sh, "pl_shader_error_diffusion: output texture must be storable");
        return false;
    }
    for (int y = 0; y <= PL_EDF_MAX_DY; y++) {
        for (int x = PL_EDF_MIN_DX; x <= PL_EDF_MAX_DX; x++) {
            // the shift mapping must move every error strictly to the right (dithering.c:305)
            if (kernel->pattern[y][x - PL_EDF_MIN_DX] && x + y * kernel->shift <= 0) {
                plb_shader_fail(sh, "invalid error diffusion kernel '%s'", kernel->name);
                return false;
            }
        }
    }
    sh->src = PLB_SRC_ERRDIFF;
    sh->is_compute = true;
    sh->out_w = width;
    sh->out_h = height;
    plb_shader_describe(sh, "error diffusion");
    plb_shader_use_tex(sh, params->input_tex);
    plb_shader_use_tex(sh, params->output_tex);
    struct plb_errdiff_launch *e = &sh->errdiff;
    memset(e, 0, sizeof(*e));
    e->src = plb_tex_priv(params->input_tex)->d;
    e->dst = plb_tex_priv(params->output_tex)->d;
    e->w = width;
    e->h = height;
    e->new_depth = params->new_depth;
    e->shift = kernel->shift;
    e->divisor = kernel->divisor;
    memcpy(e->pattern, kernel->pattern, sizeof(e->pattern));
    return true;
}

bool plb_pass_error_diffusion(pl_renderer rr, pl_dispatch dp, pl_shader *sh, int new_depth,
                              int comps, int out_w, int out_h, pl_fmt fbofmt,
                              const struct pl_error_diffusion_kernel *kernel,
                              pl_tex (*get_fbo)(void *, int, int, pl_fmt, int), void *priv)
{
    (void) rr;
    if (!fbofmt || !(fbofmt->caps & PL_FMT_CAP_STORABLE))
        return false;
    struct pl_error_diffusion_params edpars = {
        .new_depth = new_depth,
        .kernel = kernel,
    };
    edpars.input_tex = get_fbo(priv, out_w, out_h, fbofmt, comps);
    edpars.output_tex = get_fbo(priv, out_w, out_h, fbofmt, comps);
    if (!edpars.input_tex || !edpars.output_tex)
        return false;
    pl_shader dsh = pl_dispatch_begin(dp);
    if (!pl_shader_error_diffusion(dsh, &edpars)) {
        pl_dispatch_abort(dp, &dsh);
        return false;
    }
    bool ok = pl_dispatch_finish(dp, pl_dispatch_params( .shader = sh, .target = edpars.input_tex ));
    if (ok) {
        ok = pl_dispatch_compute(dp, pl_dispatch_compute_params(
            .shader = &dsh,
            .dispatch_size = {1, 1, 1},
        ));
    } else {
        pl_dispatch_abort(dp, &dsh);
    }
    *sh = pl_dispatch_begin(dp);
    pl_shader_sample_direct(*sh, pl_sample_src( .tex = ok ? edpars.output_tex : edpars.input_tex ));
    return ok;
}
