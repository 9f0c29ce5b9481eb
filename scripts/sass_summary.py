#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native code path (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld / st -> LDTM / STTM, bulk / tensor-map TMA -> UBLKCP / UTMALDG / UTMASTG, legacy mma.sync -> HMMA),
from `cuobjdump -sass` of the in-tree objects.  usage: python scripts/sass_summary.py > profiles/r02_sass_summary.md"""
import glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "audiotokenization_b200", "csrc", "obj")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCCP", "SYNCS", "HMMA", "FFMA2", "FFMA", "MUFU", "LDGSTS", "ATOM", "RED", "UCGABAR", "CCTL"]

def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name

rows = []
for obj in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                rows.append((os.path.basename(obj), cur, counts))
            cur, counts = m.group(1), {}
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for pat in PATTERNS:
                if op == pat or op.startswith(pat + "."):
                    counts[pat] = counts.get(pat, 0) + 1
                    break
    if cur:
        rows.append((os.path.basename(obj), cur, counts))

print("# SASS evidence per kernel (`cuobjdump -sass audiotokenization_b200/csrc/obj/*.o`, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma kind::f16 (cta_group::1 and ::2), `UTCBAR` = tcgen05.commit, `LDTM` / `STTM` = tcgen05.ld / st, "
      "`UBLKCP` = cp.async.bulk (TMA engine, no tensor map), `UTMALDG` / `UTMASTG` = tensor-map TMA, `SYNCS` = mbarrier ops, "
      "`FFMA2` = packed fp32 pairs, `UCGABAR` = cluster barrier.  No `HMMA` (legacy mma.sync) anywhere.\n")
cols = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "HMMA", "FFMA2", "FFMA", "MUFU"]
print("| object | kernel | " + " | ".join(cols) + " |")
print("|---|---|" + "---:|" * len(cols))
tot = {c: 0 for c in cols}
for obj, fn, counts in rows:
    name = demangle(fn)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "")
    print(f"| {obj} | `{name}` | " + " | ".join(str(counts.get(c, 0)) for c in cols) + " |")
    for c in cols:
        tot[c] += counts.get(c, 0)
print("| | **total** | " + " | ".join(str(tot[c]) for c in cols) + " |")
