#!/bin/bash
# Round profile pass (run on the GPU box through gpurun): bench line, ncu launch list of the bench,
# and one `ncu --set full` capture of each dominant kernel, summarised into gpurun_out/.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/bench_launches.csv > gpurun_out/bench_launches_agg.txt 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap chain "b200_jit_kernel|elemwise_tape_kernel" 4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train
cap gemm_bf16 "gemm_tcgen05" 2 python scripts/prof_gemm.py
cap gemm_tf32 "gemm_tcgen05" 5 python scripts/prof_gemm.py
cap reduce_row "reduce_row_fast_kernel" 6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train
cap attention_fwd "attention_fwd_kernel" 2 python train_bench.py --config lm --steps 1 --warmup 0 --eager
cap attention_bwd "attention_bwd_dq_kernel" 2 python train_bench.py --config lm --steps 1 --warmup 0 --eager
ls -la gpurun_out | tail -20
