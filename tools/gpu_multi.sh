#!/bin/bash
# multi-GPU checks (gpurun --gpus N): NCCL parity test + the bench line under torchrun
N=$(nvidia-smi -L | wc -l)
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x --timeout 500 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 5 ${BENCH_ARGS} > $O/multi_bench_n$N.json 2> $O/multi_bench_n$N.err
python - <<PY
import json
try:
    j = json.load(open("gpurun_out/multi_bench_n$N.json"))
    print("N=$N", j["value"], j["ms_per_step"], j["e2e"]["value"], [(p["pass"][:12], round(p["mean_us"])) for p in j["passes"]])
    for k, v in j.get("extra_workloads", {}).items():
        print(" ", k, v.get("value"), (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("bench failed", e, open("gpurun_out/multi_bench_n$N.err").read()[-2500:])
PY
