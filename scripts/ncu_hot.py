"""Aggregates an `ncu --page source --csv` dump: instructions executed per SASS region."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# find header row
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[h + 1:]:
    if len(r) <= iex or not r[iex]:
        continue
    try:
        data.append((r[ia], r[isrc], int(r[iex]), int(r[ismp] or 0)))
    except ValueError:
        pass
total = sum(d[2] for d in data)
print("total warp-instructions", total, "static instrs", len(data), "executed static", sum(1 for d in data if d[2]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# contiguous executed regions
regions = []
cur = None
for i, d in enumerate(data):
    if d[2] > 0:
        if cur is None:
            cur = [i, i, 0, 0]
        cur[1] = i; cur[2] += d[2]; cur[3] += d[3]
    else:
        if cur: regions.append(cur); cur = None
if cur: regions.append(cur)
regions.sort(key=lambda r: -r[2])
for r in regions[:top]:
    n = r[1] - r[0] + 1
    print(f"region {data[r[0]][0]}..{data[r[1]][0]}  static={n:5d} executed={r[2]:12d} ({100*r[2]/total:5.1f}%) samples={r[3]}")
    if len(sys.argv) > 3:
        for d in data[r[0]:r[1] + 1][: int(sys.argv[3])]:
            print("      ", d[2], d[1][:100])
