#!/bin/bash
# refresh of the round-2 evidence that does not need ncu: parity tests, smoke, bench, reference arm
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -6 > $O/r02_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err
cat $O/r02_tests.log $O/r02_smoke.log | tail -6
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench.json", "gpurun_out/r02_bench_ref.json"):
    try:
        j = json.load(open(f))
        print(f, j.get("value"), j.get("ms_per_step"), (j.get("e2e") or {}).get("value"), (j.get("cpu_baseline") or {}).get("value"),
              [(p["pass"][:12], round(p["mean_us"])) for p in j.get("passes", [])])
        for k, v in j.get("extra_workloads", {}).items():
            print("  ", k, v.get("value"), (v.get("e2e") or {}).get("value"), (v.get("parity") or {}).get("max_diff"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
