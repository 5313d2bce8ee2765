#!/bin/bash
# round 2, GPU pass 13 (1 GPU): flash attention with warp-uniform fast paths, delta from shared memory, double-buffered dS
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_flash_gpu.py tests/test_attention_gpu.py -m gpu -x -q --timeout 90 --timeout-method thread > gpurun_out/r02_pytest13.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_pytest13.log
timeout 90 python scripts/attn_bench.py 2>&1 | tail -5 | tee gpurun_out/r02_attn_bench13.txt
