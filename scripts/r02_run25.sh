#!/bin/bash
# 2 GPUs: collective stream priority high (default) vs low, graph replay and eager
set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for pr in high low; do
  B200_COLL_PRIORITY=$pr timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync fused 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/N=2 fused priority=$pr graph /"
done
B200_COLL_PRIORITY=low B200_PEER_CTAS_PER_SM=2 timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync fused 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/N=2 fused priority=low ctas=2 graph /"
B200_COLL_PRIORITY=low timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync nccl 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/N=2 nccl priority=low graph /"
