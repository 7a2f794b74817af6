#!/bin/bash
# Measurement aid (run on the GPU box through gpurun): GPU tests, then the HDR workload with
# the selectable register budgets of k_pixel, then the same with the alternative build of the
# merged EWA kernel (libplacebo_b200/alt, -DPLB_UP2_MINB=4).
export OMP_NUM_THREADS=16
timeout 700 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -8 > gpurun_out/tests2.log
for m in 3 1 2 4 5; do
  PLB_PIXEL_MINB=$m timeout 200 python bench.py --workload c3_4k_hdr10_to_sdr --steps 100 > gpurun_out/exp_c3_minb$m.json 2>/dev/null
done
timeout 100 python bench.py --workload c2_1080p_to_4k_ewa_fp16 > gpurun_out/exp_c2.json 2>/dev/null
if [ -f libplacebo_b200/alt/libplacebo_b200.so ]; then
  cp libplacebo_b200/alt/libplacebo_b200.so libplacebo_b200/libplacebo_b200.so
  for m in 3 4; do
    PLB_PIXEL_MINB=$m timeout 200 python bench.py --workload c3_4k_hdr10_to_sdr --steps 100 > gpurun_out/exp_c3_alt_minb$m.json 2>/dev/null
  done
fi
cat gpurun_out/tests2.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/exp_*.json")):
    try:
        j = json.load(open(f))
        print(f, j["value"], j["ms_per_step"], [(p["pass"][:12], round(p["mean_us"])) for p in j["passes"]])
    except Exception as e:
        print(f, "ERR", e)
PY
