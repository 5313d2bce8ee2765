#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 1 0; do
B200_FA_PERSIST=$v timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:flash_bwd_dq --csv --log-file gpurun_out/r02_dq_l2_$v.csv python scripts/attn_bench.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_dq_l2_$v.csv')) if len(r) > 10 and r[0].isdigit()]
seq = collections.defaultdict(list)
for r in rows: seq[(r[4].split('(')[0].split('::')[-1], r[-3])].append(r[-1])
for (name, metric), v in seq.items(): print('persist=$v', name[:30], metric[:34], ' '.join(v[:2]), '...', ' '.join(v[-2:]))
PY
done
