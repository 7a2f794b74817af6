#!/bin/bash
# parity tests + the full default bench line (extras, parity) — quick iteration run
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -${TAIL:-12}
timeout 900 python bench.py ${BENCH_ARGS:---no-cpu-baseline} > $O/r2b_bench.json 2> $O/r2b_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r2b_bench.json"))
    print(j["value"], j["e2e"], j.get("parity"), j.get("cpu_baseline", {}).get("value"), [(p["pass"][:12], round(p["mean_us"])) for p in j["passes"]])
    for k, v in j.get("extra_workloads", {}).items():
        print(k, v["value"], v.get("e2e", {}).get("value"), v.get("parity"), v.get("roofline", {}).get("frac"), [(p["pass"][:12], round(p["mean_us"])) for p in v.get("passes", [])])
except Exception as e:
    print("bench failed:", e, open("gpurun_out/r2b_bench.err").read()[-1500:])
PY
if [ -n "$AB" ]; then
  for v in 4 5; do
    PLB_HDR_MINB=$v timeout 300 python bench.py --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('HDR_MINB=$v', j['value'], [(p['pass'][:12], round(p['mean_us'],1)) for p in j['passes']])"
  done
fi
