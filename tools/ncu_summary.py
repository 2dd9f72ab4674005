#!/usr/bin/env python
"""Print the metrics we judge kernels by from an .ncu-rep (`ncu --set full` capture), one column per
captured launch:  python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} captured launch(es)")
    for w in WANT:
        if w in idx:
            vals = [r[idx[w]][:60] for r in rows[2:]]
            print(f"{w:<66} {units[idx[w]]:<16} " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
