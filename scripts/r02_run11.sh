#!/bin/bash
# round 2, GPU pass 11 (1 GPU): A/B of the flash backward kernels (HEAD vs previous commit) + source-level stall capture
set -u
mkdir -p gpurun_out
echo "== HEAD lib"; timeout 200 python scripts/attn_bench.py 2>&1 | tail -5 | tee gpurun_out/r02_attn_bench11_head.txt
echo "== prev lib"; BURN_B200_LIB=burn_b200/lib/libburn_b200_prevflash.so timeout 200 python scripts/attn_bench.py 2>&1 | tail -5 | tee gpurun_out/r02_attn_bench11_prev.txt
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
  ls -la gpurun_out/$name.ncu-rep
}
export BURN_B200_LIB=burn_b200/lib/libburn_b200_prevflash.so
cap r02_prev_bwd_dq "flash_bwd_dq_kernel" 2 python scripts/attn_bench.py
cap r02_prev_bwd_dkv "flash_bwd_dkv_kernel" 2 python scripts/attn_bench.py
cap r02_prev_fwd "flash_fwd_kernel" 2 python scripts/attn_bench.py
unset BURN_B200_LIB
cap r02_head_bwd_dq "flash_bwd_dq_kernel" 2 python scripts/attn_bench.py
cap r02_head_bwd_dkv "flash_bwd_dkv_kernel" 2 python scripts/attn_bench.py
