#!/bin/bash
# round 2, GPU pass 1: full parity suite, GEMM sweep (incl. graph replay), xent bench, ncu of 2048^3 / 16384^3 GEMMs
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 600 python scripts/gemm_sweep.py > gpurun_out/r02_gemm_sweep.txt 2>&1; tail -40 gpurun_out/r02_gemm_sweep.txt
timeout 120 python scripts/xent_bench.py 2>&1 | tee gpurun_out/r02_xent.txt
cap() {
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  python scripts/ncu_brief.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02_gemm2048_bf16 "gemm_tcgen05" 1 python scripts/prof_gemm2.py 2048
cap r02_gemm2048_tf32 "gemm_tcgen05" 3 python scripts/prof_gemm2.py 2048
cap r02_gemm16384_bf16 "gemm_tcgen05" 1 python scripts/prof_gemm2.py 16384
cap r02_xent "softmax_xent" 1 python scripts/xent_bench.py
grep -h "gpu__time_duration.sum\|dram__bytes\|tensor" gpurun_out/r02_gemm*.txt gpurun_out/r02_xent.txt
