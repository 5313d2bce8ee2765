#!/bin/bash
set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_bench27_n2.json 2> gpurun_out/r02_bench27_n2.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_bench27_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
t=d['train']
print({k:t.get(k) for k in ['ms_per_step','no_sync_ms_per_step','exposed_sync_ms','efficiency_vs_no_sync','sync_path','value']})
print(d['value'], d['allreduce_check'], d['peer_collective_check'])
"
tail -3 gpurun_out/r02_bench27_n2.err
