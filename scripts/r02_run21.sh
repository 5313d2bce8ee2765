#!/bin/bash
# round 2, GPU pass 21 (8 GPUs): the bench line at N=8 after the attention rewrite, LM step with NCCL sync for comparison
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r02_bench21_n8.json 2> gpurun_out/r02_bench21_n8.err; echo "bench rc=$?"; tail -1 gpurun_out/r02_bench21_n8.json | cut -c1-2500; tail -3 gpurun_out/r02_bench21_n8.err
timeout 400 $TR train_bench.py --steps 10 --warmup 3 --sync nccl 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train21_n8_nccl.txt
timeout 400 $TR train_bench.py --config encoder --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train21_n8_enc.txt
