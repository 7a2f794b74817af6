// EWA 2x phase kernels, instantiations for 1-2 components, centred phases (see k_up2.cuh)
#include "k_up2.cuh"

PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_up2a)

extern "C" int plb_up2_launch_a(pl_gpu gpu, polar_up2_args *args, int nc, int xm, int ym)
{
    if (xm == 0 && ym == 0 && nc == 1)
        return launch_up2<1, 0, 0>(gpu, *args);
    if (xm == 0 && ym == 0 && nc == 2)
        return launch_up2<2, 0, 0>(gpu, *args);
    return -1;
}
