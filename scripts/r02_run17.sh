#!/bin/bash
# round 2, GPU pass 17 (1 GPU): full GPU suite + training steps with the new flash backward kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r02_pytest17.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_pytest17.log
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train17.txt
timeout 300 python train_bench.py --config encoder --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/r02_train17_enc.txt
