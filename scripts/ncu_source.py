"""Condense `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` into the hottest CUDA source lines
(samples, share, dominant stall reasons) and the hottest SASS instructions."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
hdr = None
lines = []   # (file, line_no, source, samples, stalls dict)
sass = []    # (file, line_no, sass, samples, stalls)
cur_line = None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        stall_idx = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        s = float(r[i_samp])
    except ValueError:
        continue
    st = {}
    for h, i in stall_idx:
        try:
            v = float(r[i])
        except (ValueError, IndexError):
            v = 0
        if v:
            st[h[6:]] = v
    if r[0] != "":
        cur_line = (cur_file, r[0])
        lines.append((cur_file, r[0], r[1].strip(), s, st))
    else:
        sass.append((cur_line, r[3].strip(), s, st))
tot = sum(x[3] for x in lines) or 1
print(f"{path}: {tot:.0f} samples over {len(lines)} source lines")
def fmt(st):
    return " ".join(f"{k}={v:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
for f, ln, src, s, st in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{s:7.0f} {100*s/tot:5.1f}%  {f}:{ln:>4s}  {src[:110]}\n{'':16s}{fmt(st)}")
print("--- hottest SASS")
for cl, ins, s, st in sorted(sass, key=lambda x: -x[2])[:top]:
    print(f"{s:7.0f} {100*s/tot:5.1f}%  {cl[0]}:{cl[1]:>4s}  {ins[:90]}   [{fmt(st)}]")
