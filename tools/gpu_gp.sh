#!/bin/bash
# generic-path check: parity tests, timings of the profiling workloads, one source-level capture
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4
for wl in p1_1080p_to_1440p_ewa_generic p2_1080p_to_1440p_lanczos_ortho; do
timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$wl', j['value'], [(p['pass'][:20], round(p['mean_us'])) for p in j['passes']])"
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ortho -s 8 -c 4 -f -o /tmp/prof_p2 \
    python bench.py --workload p2_1080p_to_1440p_lanczos_ortho --steps 6 --warmup 12 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu -i /tmp/prof_p2.ncu-rep --page raw --csv > $O/r02_ncu_p2_raw.csv 2>/dev/null
ncu -i /tmp/prof_p2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r02_ncu_p2_source.csv.gz
