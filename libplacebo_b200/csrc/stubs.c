// Entry points that are declared in the public header but not implemented yet fail loudly.
#include "shader_priv.h"

bool plb_launch_ortho(pl_gpu gpu, const struct plb_ortho_launch *p)
{
    (void) p;
    PLB_ERR(gpu->log, "ortho sampling: kernel not implemented yet");
    return false;
}

bool plb_launch_errdiff(pl_gpu gpu, const struct plb_errdiff_launch *p)
{
    (void) p;
    PLB_ERR(gpu->log, "error diffusion: kernel not implemented yet");
    return false;
}

void plb_grain_obj_free(pl_gpu gpu, struct plb_grain_obj *g) { (void) gpu; (void) g; }

void plb_record_dovi_reshape(pl_shader sh, const struct pl_dovi_metadata *data)
{
    (void) data;
    plb_shader_fail(sh, "Dolby Vision reshaping: not implemented yet");
}

void plb_record_dovi_lms(pl_shader sh, const struct pl_dovi_metadata *data)
{
    (void) data;
    plb_shader_fail(sh, "Dolby Vision LMS: not implemented yet");
}

bool pl_needs_film_grain(const struct pl_film_grain_params *params) { (void) params; return false; }

bool pl_shader_film_grain(pl_shader sh, pl_shader_obj *grain_state,
                          const struct pl_film_grain_params *params)
{
    (void) grain_state; (void) params;
    plb_shader_fail(sh, "film grain: not implemented yet");
    return false;
}

size_t pl_error_diffusion_shmem_req(const struct pl_error_diffusion_kernel *kernel, int height)
{
    (void) kernel; (void) height;
    return 0;
}

bool pl_shader_error_diffusion(pl_shader sh, const struct pl_error_diffusion_params *params)
{
    (void) params;
    plb_shader_fail(sh, "error diffusion: not implemented yet");
    return false;
}
