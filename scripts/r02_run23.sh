#!/bin/bash
# 2 GPUs: overlapped vs deferred (after backward) gradient sync, fused peer kernel and NCCL
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed 's/^/N=1 /'
for mode in fused nccl; do for d in 0 1; do
  B200_SYNC_DEFER=$d timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync $mode 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/N=2 sync=$mode defer=$d /"
done; done
