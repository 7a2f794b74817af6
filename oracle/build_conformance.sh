#!/bin/bash
# TEST INFRASTRUCTURE. Builds the reference's own GPU conformance tests of the backend boundary
# (src/tests/gpu_tests.c: pl_buffer_tests, pl_texture_tests) against libplacebo_b200.so:
# gpu_tests.c is compiled where it lies under /root/reference (nothing is copied) with the
# reference's headers; unreferenced test functions (shader / GLSL tests) are dropped by
# --gc-sections, the driver is tests/c/gpu_conformance_main.c. Output: oracle/_ref/gpu_conformance
# (git-ignored, travels to the GPU box). Needs oracle/_ref/inc from build_ref.sh.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
REF=${PL_REFERENCE:-/root/reference}
W="$HERE/_ref"
[ -d "$REF/src/tests" ] && [ -d "$W/inc" ] || { echo "build_conformance: reference not available"; exit 0; }
gcc -std=c11 -O1 -D_GNU_SOURCE -DPL_HAVE_PTHREAD -DPTHREAD_HAS_SETCLOCK -DPL_HAVE_DOVI -w -fPIC \
    -ffunction-sections -fdata-sections \
    -I"$W" -I"$W/inc" -I"$REF/src/include" -I"$REF/src" -I"$REF/src/tests" -I"$REF/3rdparty/fast_float/include" \
    "$REF/src/tests/gpu_tests.c" "$ROOT/tests/c/gpu_conformance_main.c" "$ROOT/tests/c/ref_tests_shim.c" \
    -o "$W/gpu_conformance" -Wl,--gc-sections -L"$ROOT/libplacebo_b200" -lplacebo_b200 \
    -Wl,-rpath,'$ORIGIN/../../libplacebo_b200' -lm -lpthread
echo "build_conformance: OK -> $W/gpu_conformance"
