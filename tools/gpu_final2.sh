#!/bin/bash
# Round-2 evidence run (on the GPU box through gpurun): parity tests, smoke, the default bench line
# (with extras, parity and the CPU baseline), the reference arm, launch lists of every workload and
# one full ncu capture of a C3, C2, C4 and HQ-preset frame (CSV pages: the .ncu-rep files exceed
# the 64 MiB return limit).
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -6 > $O/r02_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
for w in c3:c3_4k_hdr10_to_sdr:30:12:3 c2:c2_1080p_to_4k_ewa_fp16:20:8:2 c4:c4_4k_hdr10_to_sdr_av1_grain:50:15:5 \
         hq:c3_4k_hdr10_to_sdr_hq_preset:72:12:6; do
  IFS=: read tag wl skip cnt full <<< "$w"
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
      --clock-control none -s $skip -c $cnt --csv --log-file $O/r02_launches_$tag.csv \
      python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -s $skip -c $full -f -o /tmp/prof_$tag \
      python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > $O/r02_ncu_${tag}_raw.csv 2>/dev/null
  if [ $tag = c3 ] || [ $tag = c2 ]; then
    ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r02_ncu_${tag}_source.csv.gz
  fi
done
# C5 (8K Dolby Vision + error diffusion): launch list of one frame
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -s 8 -c 8 --csv --log-file $O/r02_launches_c5.csv \
    python bench.py --workload c5_8k_dovi_errdiff --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
cat $O/r02_tests.log $O/r02_smoke.log | tail -8
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench.json", "gpurun_out/r02_bench_ref.json"):
    try:
        j = json.load(open(f))
        print(f, j.get("value"), j.get("ms_per_step"), (j.get("e2e") or {}).get("value"), j.get("cpu_baseline"),
              [(p["pass"][:12], round(p["mean_us"])) for p in j.get("passes", [])])
        for k, v in j.get("extra_workloads", {}).items():
            print("  ", k, v.get("value"), (v.get("e2e") or {}).get("value"), (v.get("parity") or {}).get("max_diff"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
