#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers the
profiles/ notes quote: duration, pipe utilisation, occupancy, stall reasons, memory traffic."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("==", r[col["Kernel Name"]][:60], r[col["Grid Size"]], r[col["Block Size"]])
        for k in KEYS:
            if k in col:
                print(f"  {k:75s} {r[col[k]]:>16s} {units[col[k]]}")
        st = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") or ("issue_stalled" in h and "_per_warp_active" in h):
                try:
                    st.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  stall {v:8.3f} {h}")


if __name__ == "__main__":
    main()
