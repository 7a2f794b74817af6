#!/bin/bash
# A/B measurement aid (run on the GPU box through gpurun): GPU parity tests with the default
# build, the render tests again with the specialised HDR kernels switched off, then the HDR
# workloads with and without them (PLB_PIXEL_HDR = 0: general interpreter kernel for the output
# pass, 3/4/5: k_pixel_hdr at that many CTAs per SM; PLB_UP2_HDR = 0: general tail in the merged
# EWA pass; PLB_PIXEL_PLANES = 0: general kernel for deband / grain plane passes, 3/4/5: CTAs/SM).
export OMP_NUM_THREADS=16
O=gpurun_out
timeout 700 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4 > $O/ab_tests.log
PLB_PIXEL_HDR=0 PLB_UP2_HDR=0 PLB_PIXEL_PLANES=0 timeout 700 python -m pytest tests/test_gpu_render.py -q -m gpu -x --timeout 300 2>&1 | tail -2 >> $O/ab_tests.log
for m in 5 4; do
  PLB_PIXEL_PLANES=$m timeout 200 python bench.py --steps 100 --no-cpu-baseline > $O/ab_c3_planes$m.json 2>/dev/null
  PLB_PIXEL_PLANES=$m timeout 200 python bench.py --workload c4_4k_hdr10_to_sdr_av1_grain --steps 100 --no-cpu-baseline > $O/ab_c4_planes$m.json 2>/dev/null
done
timeout 100 python bench.py --workload c2_1080p_to_4k_ewa_fp16 --no-cpu-baseline > $O/ab_c2.json 2>/dev/null
cat $O/ab_tests.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        j = json.load(open(f))
        print(f, j["value"], j["ms_per_step"], [(p["pass"][:12], round(p["mean_us"])) for p in j["passes"]])
    except Exception as e:
        print(f, "ERR", e)
PY
