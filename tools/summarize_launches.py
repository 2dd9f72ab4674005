#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share of device time (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((r[ix["Kernel Name"]], v))
    rows = rows[skip:]
    agg = collections.OrderedDict()
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"<.*", "", short)[-70:]
        if "eavsr" in name:
            m = re.search(r"(dcn_\w+|flow_warp_\w+|corr_\w+|conv3x3_\w+|conv_pack\w+|adapt_mix\w+|affine_offsets\w+|"
                          r"channel_sum\w+|ca_scale\w+|bias_act\w+)", name)
            short = "eavsr::" + (m.group(1) if m else short)
        c, t = agg.get(short, (0, 0.0))
        agg[short] = (c + 1, t + us)
    total = sum(t for _, t in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e3:.2f} ms of kernel time (serialised)\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {k} | {c} | {t / 1e3:.2f} | {100 * t / total:.1f}% | {t / c:.1f} |")
    ours = sum(t for k, (c, t) in agg.items() if k.startswith("eavsr::"))
    print(f"\neavsr_b200 kernels: {100 * ours / total:.1f}% of kernel time")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
