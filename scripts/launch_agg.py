"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (name + grid + block).
usage: python scripts/launch_agg.py launches.csv [skip_first_n_launches]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = {n: i for i, n in enumerate(rows[hdr])}
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
n = 0
for r in rows[hdr + 1:]:
    if r[h["Metric Name"]] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*$", "", r[h["Kernel Name"]])
    key = (name[:90], r[h["Grid Size"]], r[h["Block Size"]])
    v = float(r[h["Metric Value"]].replace(",", ""))
    unit = r[h["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    agg[key][0] += 1
    agg[key][1] += us
    total += us
byname = collections.defaultdict(lambda: [0, 0.0])
for (name, g, b), (c, t) in agg.items():
    byname[name][0] += c
    byname[name][1] += t
print(f"launches {n - skip}  total {total/1e3:.3f} ms")
print("== by kernel")
for name, (c, t) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e3:9.3f} ms {100*t/total:5.1f}%  x{c:<5d} {name}")
print("== top (kernel, grid, block)")
for (name, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t/1e3:9.3f} ms {100*t/total:5.1f}%  x{c:<5d} avg {t/c:9.1f} us  grid {g:>18s} block {b:>14s}  {name}")
