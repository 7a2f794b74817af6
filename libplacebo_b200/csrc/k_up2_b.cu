// EWA 2x phase kernels, instantiations for chroma pairs with left / top-left siting (see k_up2.cuh)
#include "k_up2.cuh"

PLB_DEFINE_PQ_UPLOAD(plb_pq_upload_up2b)

extern "C" int plb_up2_launch_b(pl_gpu gpu, polar_up2_args *args, int nc, int xm, int ym)
{
    if (nc == 2 && xm == 1 && ym == 0)
        return launch_up2<2, 1, 0>(gpu, *args);
    if (nc == 2 && xm == 1 && ym == 1)
        return launch_up2<2, 1, 1>(gpu, *args);
    return -1;
}
