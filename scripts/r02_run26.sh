#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:layer_norm_bwd_warp_kernel" -s 4 -c 1 -f -o gpurun_out/r02_lnbwd python train_bench.py --config lm --steps 1 --warmup 0 --eager > gpurun_out/r02_lnbwd.log 2>&1
ncu -i gpurun_out/r02_lnbwd.ncu-rep --page raw --csv > gpurun_out/r02_lnbwd.raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_lnbwd.raw.csv | tee gpurun_out/r02_lnbwd.txt
