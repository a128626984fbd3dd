#!/usr/bin/env python
"""Prints the handful of ncu metrics the roofline discussion uses from a .ncu-rep (run here, no GPU needed):
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== launch {k}: {name}")
        stalls = []
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:70s} {row[i]:>18s} {units[i]}")
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                stalls.append((float(row[i] or 0), h[len(STALL):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        print("  stalls (warps per issue-active):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
