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
        print(k, v["value"], v["e2e"]["value"], v.get("parity"), v["roofline"]["frac"], [(p["pass"][:12], round(p["mean_us"])) for p in v["passes"]])
except Exception as e:
    print("bench failed:", e, open("gpurun_out/r2b_bench.err").read()[-1500:])
PY
