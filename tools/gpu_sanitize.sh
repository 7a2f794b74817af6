#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory kernels) over a subset of the
# parity tests that covers every kernel family.
O=gpurun_out
mkdir -p $O
SEL='test_c3_hdr10_to_sdr or test_c2_sdr_upscale_2x_ewa_lanczos or planar_yuv_target or custom_lut or error_diffusion_depths or contrast_recovery or c5_dolby or test_render_with_av1_grain or test_polar_generic or ortho_upscale_2x or mirror_separable or blit'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck.log \
    python -m pytest tests -q -m gpu -x --timeout 1200 -k "$SEL" 2>&1 | tail -4
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" $O/memcheck.log; tail -3 $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/racecheck.log \
    python -m pytest tests -q -m gpu -x --timeout 800 -k "error_diffusion_depths or test_c2_sdr_upscale_2x_ewa_lanczos" 2>&1 | tail -3
echo "racecheck rc=$?"; grep -c "hazard" $O/racecheck.log; tail -3 $O/racecheck.log
