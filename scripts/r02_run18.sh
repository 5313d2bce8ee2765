#!/bin/bash
# launch lists (per-kernel device time, cold-cache / serialised) of one encoder step and one LM step
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_train_enc_launches.csv \
    python train_bench.py --config encoder --steps 1 --warmup 0 --eager > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_train_enc_launches.csv > gpurun_out/r02_train_enc_agg.txt 2>&1; sed -n 1,45p gpurun_out/r02_train_enc_agg.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_train_lm_launches.csv \
    python train_bench.py --config lm --steps 1 --warmup 0 --eager > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_train_lm_launches.csv > gpurun_out/r02_train_lm_agg.txt 2>&1; sed -n 1,22p gpurun_out/r02_train_lm_agg.txt
