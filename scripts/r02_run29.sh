#!/bin/bash
# the driver's own 8-GPU command, once more, on the final tree
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r02_bench29_n8.json 2> gpurun_out/r02_bench29_n8.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_bench29_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
t=d['train']
print({k:t.get(k) for k in ['ms_per_step','no_sync_ms_per_step','exposed_sync_ms','efficiency_vs_no_sync','value','replica_checksum_spread']})
print(d['value'], d['e2e']['value'], d['allreduce_check'], d['peer_collective_check'], t['encoder_configs3']['ms_per_step'])
"
