#!/bin/bash
# one full ncu capture of the cluster error-diffusion kernel (8K frame)
O=gpurun_out; mkdir -p $O
timeout 400 ncu --set full --clock-control none -k regex:k_errdiff_cluster -s 1 -c 1 -f -o /tmp/prof_c5 \
    python bench.py --workload c5_8k_dovi_errdiff --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu -i /tmp/prof_c5.ncu-rep --page raw --csv > $O/r02_ncu_c5_errdiff_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02_ncu_c5_errdiff_raw.csv 2>&1 | head -40
