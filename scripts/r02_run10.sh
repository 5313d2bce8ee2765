#!/bin/bash
# round 2, GPU pass 10 (1 GPU): double-buffered flash backward, short-axis column reduce, one LN-backward finisher
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method thread > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_pytest10.log
timeout 200 python scripts/attn_bench.py 2>&1 | tail -5 | tee gpurun_out/r02_attn_bench10.txt
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train10.txt
timeout 300 python train_bench.py --config encoder --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train10_enc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_train_lm_launches.csv \
    python train_bench.py --config lm --steps 1 --warmup 0 --eager > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_train_lm_launches.csv > gpurun_out/r02_train_lm_agg.txt 2>&1; sed -n 1,22p gpurun_out/r02_train_lm_agg.txt
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02_flash_bwd_dq "flash_bwd_dq_kernel" 2 python scripts/attn_bench.py
cap r02_flash_bwd_dkv "flash_bwd_dkv_kernel" 2 python scripts/attn_bench.py
