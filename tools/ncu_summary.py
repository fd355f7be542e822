#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed):  python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "smsp__inst_executed_op_tma_ld.sum",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_sleeping", "smsp__pcsamp_warps_issue_stalled_membar",
]


def main():
    rep = sys.argv[1]
    npts = float(sys.argv[2]) if len(sys.argv) > 2 else 512.0 ** 3
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    for r in data:
        print("kernel:", r[kn][:110])
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = (r[i], units[i])
                print(f"  {w:82s} {r[i]:>16s} {units[i]}")
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(vals["dram__bytes_read.sum"][0]) * scale[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * scale[vals["dram__bytes_write.sum"][1]]
            tus = float(vals["gpu__time_duration.sum"][0]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[vals["gpu__time_duration.sum"][1]]
            inst = float(vals["smsp__inst_executed.sum"][0])
            print(f"  -> DRAM traffic {rd + wr:.4g} B = {(rd + wr) / npts:.2f} B/pt (algorithmic 16), "
                  f"{(rd + wr) / tus / 1e3:.0f} GB/s under ncu, {npts / tus / 1e3:.1f} Gpts/s under ncu, "
                  f"{inst * 32 / npts:.1f} lane-instr/pt")
        except Exception as e:  # noqa: BLE001
            print("  (derived numbers unavailable:", e, ")")


if __name__ == "__main__":
    main()
