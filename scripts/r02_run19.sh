#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python scripts/fused_reduce_bench.py | tee gpurun_out/r02_fused_reduce19.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:b200_jit_cols -s 2 -c 1 -f -o gpurun_out/r02_fused_cols python scripts/fused_reduce_bench.py > gpurun_out/r02_fused_cols.log 2>&1
ncu -i gpurun_out/r02_fused_cols.ncu-rep --page raw --csv > gpurun_out/r02_fused_cols.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_fused_cols.raw.csv | tee gpurun_out/r02_fused_cols.txt | head -40
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 python scripts/fused_reduce_bench.py 2>&1 | grep -E "b200_jit|reduce_|gpu__time" | head -24
