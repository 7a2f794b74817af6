#!/bin/bash
# launch lists of the kernel families outside the BASELINE configs (generic polar, separable)
O=gpurun_out
mkdir -p $O
for w in p1:p1_1080p_to_1440p_ewa_generic p2:p2_1080p_to_1440p_lanczos_ortho; do
  tag=${w%%:*}; wl=${w##*:}
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-extra > $O/r02_bench_$tag.json 2> $O/r02_bench_$tag.err
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
      --clock-control none -s 30 -c 10 --csv --log-file $O/r02_launches_$tag.csv \
      python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
  python - <<PY
import json, csv
try:
    j = json.load(open("gpurun_out/r02_bench_$tag.json"))
    print("$tag", j["value"], [(p["pass"][:24], round(p["mean_us"])) for p in j["passes"]])
except Exception as e:
    print("$tag bench failed", e, open("gpurun_out/r02_bench_$tag.err").read()[-1200:])
rows = list(csv.reader(open("gpurun_out/r02_launches_$tag.csv")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r); H = rows[h]
kn, mn, mv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
for r in rows[h + 1:h + 41]:
    if len(r) > mv and r[mn] == "gpu__time_duration.sum":
        print("  ", r[kn][:70], r[mv])
PY
done
