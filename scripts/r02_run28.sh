#!/bin/bash
set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r02_bench28_n4.json 2> gpurun_out/r02_bench28_n4.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_bench28_n4.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
t=d['train']
print({k:t.get(k) for k in ['ms_per_step','no_sync_ms_per_step','exposed_sync_ms','efficiency_vs_no_sync','value','replica_checksum_spread']})
print(d['value'], d['e2e']['value'], d['allreduce_check'], d['peer_collective_check'], t['encoder_configs3']['ms_per_step'])
"
tail -2 gpurun_out/r02_bench28_n4.err
