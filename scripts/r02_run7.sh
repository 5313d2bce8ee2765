#!/bin/bash
# round 2, GPU pass 7 (2 GPUs): peer kernel with 8 loads in flight per thread; grid sweep; LM step per sync mode
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for c in 1 2 4; do
  B200_PEER_CTAS_PER_SM=$c timeout 300 $TR scripts/peer_bench.py 2>&1 | grep bucket_mib | sed "s/^/ctas_per_sm=$c /" | tee -a gpurun_out/r02_peer_bench_n2b.txt
done
for mode in nccl fused; do
  timeout 600 $TR train_bench.py --steps 10 --warmup 3 --sync $mode > gpurun_out/r02_train_n2b_$mode.txt 2>&1; echo "train $mode rc=$?"; tail -1 gpurun_out/r02_train_n2b_$mode.txt | cut -c1-300
done
B200_PEER_CTAS_PER_SM=1 timeout 600 $TR train_bench.py --steps 10 --warmup 3 --sync fused 2>&1 | tail -1 | cut -c1-300
timeout 300 python train_bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m pytest tests/test_jit_gpu.py tests/test_elemwise_gpu.py tests/test_collective_gpu.py -m gpu -x -q --timeout 300 --timeout-method thread 2>&1 | tail -5
