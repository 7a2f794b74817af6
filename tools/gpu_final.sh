#!/bin/bash
# Round-end measurement run (on the GPU box through gpurun): parity tests, smoke, the bench lines,
# the reference arm, and the ncu evidence (launch lists + one full capture per workload, exported
# as CSV pages because the .ncu-rep files exceed the 64 MiB return limit).
export OMP_NUM_THREADS=16
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -8 > $O/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1
timeout 300 python bench.py > $O/final_bench_c3.json 2> $O/final_bench_c3.err
timeout 300 python bench.py --workload c2_1080p_to_4k_ewa_fp16 > $O/final_bench_c2.json 2> $O/final_bench_c2.err
timeout 300 python bench.py --workload c4_4k_hdr10_to_sdr_av1_grain > $O/final_bench_c4.json 2> $O/final_bench_c4.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err
for w in c3:c3_4k_hdr10_to_sdr c2:c2_1080p_to_4k_ewa_fp16; do
  tag=${w%%:*}; wl=${w##*:}
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 12 \
      --csv --log-file $O/final_launches_$tag.csv python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline > /dev/null 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on -s 30 -c 3 -f -o /tmp/prof_$tag \
      python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline > /dev/null 2>&1
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $O/final_ncu_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/final_ncu_${tag}_source.csv.gz
done
# 2-GPU line (only when the box has two GPUs: gpurun --gpus 2)
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 100 --warmup 5 > $O/final_bench_c3_2gpu.json 2> $O/final_bench_c3_2gpu.err
fi
cat $O/final_tests.log $O/final_smoke.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*.json")):
    try:
        j = json.load(open(f))
        print(f, j.get("value"), j.get("ms_per_step"), (j.get("e2e") or {}).get("value"),
              [(p["pass"][:12], round(p["mean_us"])) for p in j.get("passes", [])])
    except Exception as e:
        print(f, "ERR", e)
PY
