#!/bin/bash
# per-launch device times of the flash kernels (cold-cache, serialised) + one full capture of the dQ kernel
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:flash --csv --log-file gpurun_out/r02_attn_launches16.csv python scripts/attn_bench.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_attn_launches16.csv')) if len(r) > 10 and r[0].isdigit()]
seq = collections.defaultdict(list)
for r in rows:
    name = r[4].split('(')[0].split('::')[-1]
    if 'time_duration' in r[-3] or 'gpu__time' in r[12] if len(r) > 12 else False: pass
for r in rows:
    name = r[4].split('(')[0].split('::')[-1]; metric = r[-3]; val = r[-1]
    seq[(name, metric)].append(val)
for (name, metric), v in seq.items():
    print(name, metric.split('.')[0], ' '.join(v[:3]), '...', ' '.join(v[-3:]))
PY
