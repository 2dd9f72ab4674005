#!/usr/bin/env python
"""Top stall sites of a kernel from an ncu report (source page, SASS view).
usage: python tools/ncu_stalls.py report.ncu-rep [launch_index] [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
lo = heads[which]
hi = heads[which + 1] if which + 1 < len(heads) else len(rows)
h = rows[lo]
si, ai, ni = h.index("Source"), h.index("Address"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
body = [r for r in rows[lo + 1:hi] if len(r) == len(h)]
tot = sum(int(r[ni] or 0) for r in body)
print(f"launch {which}: {len(body)} instructions, {tot} samples")
agg = {}
for r in body:
    for i, c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[i] or 0)
print("stall totals:", ", ".join(f"{c[6:]}={v}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
for idx, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][ni] or 0))[:top]:
    why = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {int(r[ni]):6d} {100.0 * int(r[ni]) / max(tot, 1):5.1f}%  {r[si][:90]:90s} {why}")
