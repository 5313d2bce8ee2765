"""Opcode evidence per cubin of libburn_b200.so: counts of the Blackwell-only SASS mnemonics (tcgen05 = UTC*MMA / LDTM /
UTCBAR, TMA = UTMALDG / UTMASTG / UBLKCP, cp.async = LDGSTS, clusters = UCGABAR / CCTL…), per kernel family.
Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt   (runs on the CPU build box: cuobjdump only)"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "burn_b200" / "lib" / "libburn_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF",
        "LDGSTS", "SYNCS", "UCGABAR", "MUFU.EX2", "HMMA", "IMMA", "DMMA", "RED", "ATOM", "LDG.E.128", "STG.E.128", "LD.E", "ST.E"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
fn, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[fn]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[fn][k] += 1
                if k in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM") and ("2CTA" in op or "MULTICAST" in op):
                    per[fn][k + " (.2CTA/.MULTICAST)"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode summary of {LIB.relative_to(ROOT)} (cuobjdump -sass, sm_100a); {len(per)} kernels")
fam = collections.OrderedDict()
for raw, name in zip(per, demangle):
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"<.*", "", short).split("::")[-1]
    f = fam.setdefault(short, {"n": 0, "c": collections.Counter()})
    f["n"] += 1
    f["c"] += per[raw]
for short, f in sorted(fam.items(), key=lambda kv: -kv[1]["c"]["_total"]):
    c = f["c"]
    hits = ", ".join(f"{k} {c[k]}" for k in c if k != "_total" and c[k])
    print(f"{short:44s} x{f['n']:<4d} {c['_total']:8d} instr   {hits}")
