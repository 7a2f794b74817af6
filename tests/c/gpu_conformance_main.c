/* TEST INFRASTRUCTURE — driver for the reference's own GPU conformance tests of the backend
 * boundary (src/tests/gpu_tests.c: pl_buffer_tests :10, pl_texture_tests :226), compiled in
 * place from /root/reference (nothing is copied) and linked against libplacebo_b200.so. The
 * sibling of the reference's per-backend mains (src/tests/vulkan.c, opengl_surfaceless.c,
 * dummy.c), with pl_cuda_create as the constructor. Built by oracle/build_conformance.sh into
 * oracle/_ref/gpu_conformance (git-ignored; travels to the GPU box like the other built
 * files); run by tests/test_gpu_conformance.py. */
#include "gpu_tests.h"

/* prototype only: the test is compiled against the reference's headers, which do not carry the
 * product's constructor (include/libplacebo_b200.h:pl_cuda_create) */
struct pl_cuda_params_min { int device; void *stream; int no_async; };
pl_gpu pl_cuda_create(pl_log log, const void *params);
void pl_cuda_destroy(pl_gpu *gpu);

int main(void)
{
    pl_log log = pl_test_logger();
    pl_gpu gpu = pl_cuda_create(log, NULL);
    if (!gpu) {
        printf("no CUDA device: skipping\n");
        pl_log_destroy(&log);
        return SKIP;
    }
    pl_buffer_tests(gpu);
    pl_texture_tests(gpu);
    pl_cuda_destroy(&gpu);
    pl_log_destroy(&log);
    printf("conformance: ok\n");
    return 0;
}
