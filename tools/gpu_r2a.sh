#!/bin/bash
# Round-2 iteration run: parity tests, bench lines, launch list + full ncu captures of C3 and C2.
export OMP_NUM_THREADS=16
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -15 > $O/r2a_tests.log
timeout 300 python bench.py --no-cpu-baseline > $O/r2a_bench_c3.json 2> $O/r2a_bench_c3.err
timeout 300 python bench.py --workload c2_1080p_to_4k_ewa_fp16 --no-cpu-baseline > $O/r2a_bench_c2.json 2> $O/r2a_bench_c2.err
for w in c3:c3_4k_hdr10_to_sdr c2:c2_1080p_to_4k_ewa_fp16; do
  tag=${w%%:*}; wl=${w##*:}
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 30 -c 12 \
      --csv --log-file $O/r2a_launches_$tag.csv python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on -s 30 -c 3 -f -o /tmp/prof_$tag \
      python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $O/r2a_ncu_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r2a_ncu_${tag}_source.csv.gz
done
cat $O/r2a_tests.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_bench_*.json")):
    try:
        j = json.load(open(f))
        print(f, j.get("value"), j.get("ms_per_step"), (j.get("e2e") or {}).get("value"),
              [(p["pass"][:12], round(p["mean_us"])) for p in j.get("passes", [])])
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-800:])
PY
