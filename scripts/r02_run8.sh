#!/bin/bash
# round 2, GPU pass 8 (8 GPUs): peer-memory collectives at full width — value checks + bucket timings, LM step nccl vs fused
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/peer_bench.py 2>&1 | grep -E "bucket_mib|ok|Error|error" | tee gpurun_out/r02_peer_bench_n8.txt
for mode in fused nccl; do
  timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync $mode > gpurun_out/r02_train_n8_$mode.txt 2>&1; echo "train $mode rc=$?"; tail -1 gpurun_out/r02_train_n8_$mode.txt | cut -c1-260
done
timeout 400 $TR train_bench.py --config encoder --steps 10 --warmup 3 --sync fused 2>&1 | tail -1 | cut -c1-260 | tee gpurun_out/r02_train_n8_enc_fused.txt
