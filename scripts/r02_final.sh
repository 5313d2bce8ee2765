#!/bin/bash
# round 2 final profile pass (1 GPU): bench line, ncu launch list of the same command, one full capture of the chain kernel,
# micro-benches, launch lists of the two training steps, brief captures of the three flash kernels
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench_final.json | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_bench_launches.csv > gpurun_out/r02_bench_launches_agg.txt 2>&1; head -14 gpurun_out/r02_bench_launches_agg.txt
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  python scripts/ncu_brief.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02_chain "b200_jit_kernel" 4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train
timeout 120 python scripts/reduce_bench.py 2>&1 | tee gpurun_out/r02_reduce_bench.txt
timeout 120 python scripts/fused_reduce_bench.py 2>&1 | tee gpurun_out/r02_fused_reduce.txt
timeout 120 python scripts/xent_bench.py 2>&1 | tee gpurun_out/r02_xent_final.txt
timeout 120 python scripts/quick_bench.py 2>&1 | tee gpurun_out/r02_quick_bench.txt
timeout 200 python scripts/attn_bench.py 2>&1 | tee gpurun_out/r02_attn_bench_final.txt
timeout 300 python scripts/gemm_sweep.py 2>&1 > gpurun_out/r02_gemm_sweep_final.txt; tail -3 gpurun_out/r02_gemm_sweep_final.txt
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train_final.txt
timeout 300 python train_bench.py --config encoder --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train_final_enc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_train_lm_launches.csv \
    python train_bench.py --config lm --steps 1 --warmup 0 --eager > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_train_lm_launches.csv > gpurun_out/r02_train_lm_agg.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_train_enc_launches.csv \
    python train_bench.py --config encoder --steps 1 --warmup 0 --eager > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_train_enc_launches.csv > gpurun_out/r02_train_enc_agg.txt 2>&1
cap r02_flash_fwd "flash_fwd_kernel" 2 python scripts/attn_bench.py
cap r02_flash_bwd_dq "flash_bwd_dq_kernel" 2 python scripts/attn_bench.py
cap r02_flash_bwd_dkv "flash_bwd_dkv_kernel" 2 python scripts/attn_bench.py
