#!/bin/bash
# generic-path check: parity tests + timings of the profiling workloads
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -4
for wl in p1_1080p_to_1440p_ewa_generic p2_1080p_to_1440p_lanczos_ortho; do
timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$wl', j['value'], [(p['pass'][:20], round(p['mean_us'])) for p in j['passes']])"
done
