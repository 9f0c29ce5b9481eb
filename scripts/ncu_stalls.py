#!/usr/bin/env python
"""Source-level stall summary of one profiled launch: ncu_stalls.py <report.ncu-rep> <launch index> [top N]
(warp-state samples per SASS instruction from `ncu --set full --import-source on`; idle warps parked at the final barrier
or in NANOSLEEP are listed separately so that the working warps' stalls can be read directly)."""
import csv, subprocess, sys, io

rep, launch = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
name = rows[0][1][:90]
hdr = rows[1]
def num(x):
    try: return float(x)
    except ValueError: return 0.0
body = [r for r in rows[2:] if len(r) > 45 and r[0].startswith("0x")]
# ncu lists the function once per source view; drop an exact second copy
half = len(body) // 2
if half and all(body[i][1] == body[i + half][1] for i in range(0, half, max(1, half // 50))):
    body = body[:half]
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
idle = lambda src: ("NANOSLEEP" in src) or ("MEMBAR" in src) or ("EXIT" in src) or ("BAR.SYNC" in src) or ("WARPSYNC" in src) or ("UCGABAR" in src)
tot_all = sum(num(r[2]) for r in body)
tot_idle = sum(num(r[2]) for r in body if idle(r[1]))
print(f"== {name}\n   {tot_all:.0f} warp samples, {100 * tot_idle / tot_all:.0f} % parked (sleeping on a barrier / final barrier / exit)")
work = [r for r in body if not idle(r[1])]
tw = sum(num(r[2]) for r in work)
by = sorted(((sum(num(r[i]) for r in work), h) for i, h in reasons), reverse=True)
print("   working warps: " + "  ".join(f"{h[6:]} {100 * v / tw:.0f}%" for v, h in by[:9]))
for v, k, src in sorted(((num(r[2]), k, r[1].strip()) for k, r in enumerate(body) if not idle(r[1])), reverse=True)[:top_n]:
    i_best = max(reasons, key=lambda ih: num(body[k][ih[0]]))
    print(f"   {100 * v / tw:5.1f}%  #{k:5d}  {i_best[1][6:]:12s} {src[:84]}")
