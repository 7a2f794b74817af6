#!/bin/bash
# launch list + full ncu capture of the HQ-preset workload's passes
O=gpurun_out
mkdir -p $O
wl=c3_4k_hdr10_to_sdr_hq_preset
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 72 -c 12 \
    --csv --log-file $O/hq_launches.csv python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ortho -s 8 -c 2 -f -o /tmp/prof_hq \
    python bench.py --workload $wl --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu -i /tmp/prof_hq.ncu-rep --page raw --csv > $O/hq_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_hq.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/hq_ncu_source.csv.gz
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/hq_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
for r in rows[hdr + 1:]:
    if len(r) > mv and r[mn] in ("gpu__time_duration.sum", "smsp__inst_executed.sum"):
        print(r[kn][:70], r[mn][:20], r[mv])
PY
