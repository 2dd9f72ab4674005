#!/usr/bin/env python
"""Time the DCN forward (bench shape, graph replay) for every library variant in build/abl/ (one subprocess per
variant: EAVSR_B200_LIB selects the .so).  Timing only -- ablated variants compute wrong results on purpose."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch
import bench
dev = torch.device("cuda:0")
peaks = bench.load_peaks()
with torch.no_grad():
    r = bench.dcn_roofline(dev, peaks, dg=8, iters=40, nested=False)
    out = {"us": r["us_per_launch"]}
    if os.environ.get("ABL_AFF", "1") == "1":
        out["aff_us"] = bench.dcn_affine_us(dev, peaks, iters=40)["us_per_launch"]
print("RESULT " + json.dumps(out))
''' % ROOT
libs = sorted(f for f in os.listdir(os.path.join(ROOT, "build", "abl")) if f.endswith(".so"))
res = {}
for lib in libs:
    env = dict(os.environ, EAVSR_B200_LIB=os.path.join(ROOT, "build", "abl", lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    res[lib] = json.loads(line[0][7:]) if line else {"failed": (r.stdout + r.stderr)[-400:]}
    print(lib, res[lib], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "abl.json"), "w"), indent=1)
