#!/bin/bash
set -u
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
}
cap r02_v20_bwd_dkv "flash_bwd_dkv_kernel" 2 python scripts/attn_bench.py
cap r02_v20_bwd_dq "flash_bwd_dq_kernel" 2 python scripts/attn_bench.py
cap r02_v20_fwd "flash_fwd_kernel" 2 python scripts/attn_bench.py
