// EWA 2x phase kernels, instantiations for 3-4 components, centred phases (see k_up2.cuh)
#include "k_up2.cuh"

PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_up2c)

extern "C" int plb_up2_launch_c(pl_gpu gpu, polar_up2_args *args, int nc, int xm, int ym)
{
    if (xm == 0 && ym == 0 && nc == 3)
        return launch_up2<3, 0, 0>(gpu, *args);
    if (xm == 0 && ym == 0 && nc == 4)
        return launch_up2<4, 0, 0>(gpu, *args);
    return -1;
}
