#!/bin/bash
# round 2, GPU pass 6 (2 GPUs): peer-memory collectives — single-process two-device test, multi-process value checks and
# bucket timings, the LM step under the three gradient-sync modes, the bench line at N=2
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_collective_gpu.py tests/test_fusion_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest6.log
timeout 600 $TR scripts/peer_bench.py > gpurun_out/r02_peer_bench_n2.txt 2>&1; echo "peer_bench rc=$?"; tail -12 gpurun_out/r02_peer_bench_n2.txt
for mode in nccl peer fused; do
  timeout 600 $TR train_bench.py --steps 10 --warmup 3 --sync $mode > gpurun_out/r02_train_n2_$mode.txt 2>&1; echo "train $mode rc=$?"; tail -3 gpurun_out/r02_train_n2_$mode.txt | cut -c1-1500
done
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench6_n2.json 2> gpurun_out/r02_bench6_n2.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02_bench6_n2.json; tail -5 gpurun_out/r02_bench6_n2.err
timeout 300 python train_bench.py --via-stream --steps 20 --warmup 3 > gpurun_out/r02_via_stream.json 2> gpurun_out/r02_via_stream.err; echo "via-stream rc=$?"; cat gpurun_out/r02_via_stream.json; tail -3 gpurun_out/r02_via_stream.err
