#!/bin/bash
# round 2, GPU pass 15 (1 GPU): suspend-hinted mbarrier waits (attention + GEMM), tensor-memory dS in the dQ kernel
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_attention_flash_gpu.py tests/test_attention_gpu.py tests/test_matmul_gpu.py -m gpu -x -q --timeout 120 --timeout-method thread > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest15.log
timeout 90 python scripts/attn_bench.py 2>&1 | tail -5 | tee gpurun_out/r02_attn_bench15.txt
timeout 300 python scripts/gemm_sweep.py 2>&1 | tee gpurun_out/r02_gemm_sweep15.txt | head -24
