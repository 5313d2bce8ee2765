#!/bin/bash
# round 2, GPU pass 2: flash attention parity (first), then the rest of the suite, xent bench, attention bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_flash_gpu.py -x -q --timeout 120 --timeout-method thread > gpurun_out/r02_flash_test.log 2>&1; echo "flash rc=$?"
tail -25 gpurun_out/r02_flash_test.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_attention_flash_gpu.py > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_pytest2.log
timeout 120 python scripts/xent_bench.py 2>&1 | tee gpurun_out/r02_xent2.txt
timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/r02_attn_bench.txt
