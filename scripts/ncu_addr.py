"""Bucket the samples of an `ncu --page source --csv --print-source cuda,sass` export by SASS address range
(0x800-byte buckets) with the attention_flash.cu source lines each bucket covers: separates the warp roles
(producer / issuer / row threads) even when their waits are the same inlined helper."""
import csv, sys, collections
path = sys.argv[1]
hdr = None; cur = None; curfile = None
seen = {}
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": curfile = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; i_s = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "":
        try: cur = (curfile, int(r[0]))
        except ValueError: pass
        continue
    try: s = float(r[i_s])
    except ValueError: continue
    a = r[2]
    if a != '...' and a not in seen: seen[a] = (cur, r[3].strip(), s)
addrs = sorted(seen); base = int(addrs[0], 16)
tot = sum(v[2] for v in seen.values())
print('total samples', tot)
buck = collections.OrderedDict()
for a in addrs:
    c, ins, s = seen[a]; off = int(a, 16) - base
    b = buck.setdefault(off // 0x800, [0, set(), 0])
    b[0] += s
    if c and c[0].endswith('.cu'): b[1].add(c[1])
    if 'UTCHMMA' in ins or 'UTMALDG' in ins: b[2] += 1
for k, (s, lines, n) in buck.items():
    if s > 0: print(f"{k*0x800:6x} {s:6.0f} {100*s/tot:5.1f}%  lines {min(lines) if lines else ''}-{max(lines) if lines else ''}  mma/tma instrs {n}")
