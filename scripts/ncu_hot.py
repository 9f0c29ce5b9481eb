#!/usr/bin/env python
"""Top SASS instructions by executed count / stall samples for one profiled launch.
usage: ncu_hot.py <rep> <launch-index> [top]"""
import csv, io, subprocess, sys, collections
rep, idx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{idx}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if r.get("Instructions Executed") not in (None, "") and r["Instructions Executed"].isdigit()]
tot_inst = sum(int(r["Instructions Executed"]) for r in rows)
tot_samp = sum(int(r["# Samples"]) for r in rows)
print(f"{len(rows)} SASS instructions, {tot_inst} warp-instructions executed, {tot_samp} samples")
op = collections.Counter(); ops = collections.Counter()
for r in rows:
    m = r["Source"].split()
    name = m[1] if m and m[0].startswith("@") else (m[0] if m else "?")
    name = name.split(".")[0]
    op[name] += int(r["Instructions Executed"]); ops[name] += int(r["# Samples"])
print("\nby opcode (executed warp-instr, share; stall samples share):")
for k, v in op.most_common(28):
    print(f"  {k:12s} {v:12d} {100*v/tot_inst:5.1f}%   samples {100*ops[k]/max(tot_samp,1):5.1f}%")
print("\ntop instructions by stall samples:")
for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:top]:
    print(f"  {int(r['# Samples']):7d} {int(r['Instructions Executed']):10d}  {r['Source'].strip()[:90]}")
