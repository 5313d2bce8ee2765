#!/bin/bash
# 2 GPUs: fused peer sync with 1 / 2 / 4 sync CTAs per SM launched (one fits beside a persistent GEMM CTA)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for c in 1 2 4; do
  B200_PEER_CTAS_PER_SM=$c timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync fused 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/N=2 fused ctas_per_sm=$c /"
done
B200_PEER_CTAS_PER_SM=1 timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync fused --bucket-mb 16 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/N=2 fused ctas_per_sm=1 bucket 16MiB /"
