#!/bin/bash
# round 2, GPU pass 5: the rewritten host fusion layer (relative plans, fuser/optimization handles, views, in place),
# the whole suite, the via-stream encoder forward
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r02_pytest5a.log 2>&1; echo "fusion rc=$?"
tail -30 gpurun_out/r02_pytest5a.log
timeout 600 python train_bench.py --via-stream --steps 20 --warmup 3 > gpurun_out/r02_via_stream.json 2> gpurun_out/r02_via_stream.err; echo "via-stream rc=$?"; cat gpurun_out/r02_via_stream.json; tail -5 gpurun_out/r02_via_stream.err
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest5.log
