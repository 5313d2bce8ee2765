#!/bin/bash
# round 2, GPU pass 4: new tests (loss curves, flash fixes, movement ops), then the bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_loss_curve_gpu.py tests/test_attention_flash_gpu.py tests/test_index_gpu.py tests/test_golden_gpu.py -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_pytest4.log
timeout 600 python bench.py > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench rc=$?"; cat gpurun_out/r02_bench4.json; tail -5 gpurun_out/r02_bench4.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 | tee gpurun_out/r02_bench4_ref.json
