#!/bin/bash
# round 2, GPU pass 9 (1 GPU): f32 tanh / cube in the gelu backward, erf region skip, packed-key arg reduce, cluster xent
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method thread > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_pytest9.log
timeout 120 python scripts/reduce_bench.py 2>&1 | tee gpurun_out/r02_reduce_bench.txt
timeout 120 python scripts/fused_reduce_bench.py 2>&1 | tee gpurun_out/r02_fused_reduce.txt
timeout 120 python scripts/xent_bench.py 2>&1 | tee gpurun_out/r02_xent3.txt
B200_XENT_NO_PAIR=1 timeout 120 python scripts/xent_bench.py 2>&1 | sed 's/^/no-pair: /' | tee -a gpurun_out/r02_xent3.txt
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train9.txt
timeout 300 python train_bench.py --config encoder --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train9_enc.txt
timeout 600 python bench.py > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench9.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_train_lm_launches.csv \
    python train_bench.py --config lm --steps 1 --warmup 0 --eager > /dev/null 2>&1
python scripts/launch_agg.py gpurun_out/r02_train_lm_launches.csv > gpurun_out/r02_train_lm_agg.txt 2>&1; head -24 gpurun_out/r02_train_lm_agg.txt
cap() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$name.raw.csv > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02_argmax "reduce_row_fast_kernel" 250 python scripts/reduce_bench.py
cap r02_fused_reduce_rows "b200_jit_rows" 2 python scripts/fused_reduce_bench.py
cap r02_xent_k "softmax_xent_pair_kernel" 1 python scripts/xent_bench.py
