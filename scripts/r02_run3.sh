#!/bin/bash
# round 2, GPU pass 3: whole suite (flash attention + expression goldens + movement ops), bench line, attention bench
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02_pytest3.log
timeout 300 python bench.py > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"; cat gpurun_out/r02_bench3.json
timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/r02_attn_bench3.txt
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/r02_train3.txt
