cap() {
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap attention_fwd "attention_fwd_kernel" 2 python train_bench.py --config lm --steps 1 --warmup 0 --eager
cap attention_bwd "attention_bwd_dq_kernel" 2 python train_bench.py --config lm --steps 1 --warmup 0 --eager
