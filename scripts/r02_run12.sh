#!/bin/bash
# round 2, GPU pass 12 (8 GPUs): the bench line at N=8 (value checks, fused peer gradient sync), LM step with NCCL sync for comparison
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r02_bench12_n8.json 2> gpurun_out/r02_bench12_n8.err; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench12_n8.json | cut -c1-3000; tail -5 gpurun_out/r02_bench12_n8.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync nccl > gpurun_out/r02_train_n8_nccl.txt 2>&1; echo "train nccl rc=$?"; tail -1 gpurun_out/r02_train_n8_nccl.txt | cut -c1-600
