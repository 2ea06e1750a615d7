#!/usr/bin/env python
"""Condense ncu reports (gpurun_out/*.ncu-rep) into the small text summaries kept under profiles/.
Usage: python tools/ncu_summary.py <report.ncu-rep> [units_per_launch] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_global_st.sum",
]


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, unit = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none, report {rep.split('/')[-1]}")
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        u = dict(zip(hdr, unit))
        print(f"\nkernel: {rec.get('Kernel Name')}")
        for k in KEYS:
            if k in rec and rec[k] != "":
                print(f"  {k:86s} {rec[k]:>16s} {u[k]}")
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd = float(rec["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
            wr = float(rec["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
            tscale = {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
            t = float(rec["gpu__time_duration.sum"]) * tscale[u["gpu__time_duration.sum"]]
            print(f"  derived: dram traffic {rd + wr:.4e} B per launch = {(rd + wr) / t / 1e9:.0f} GB/s over {t * 1e3:.3f} ms")
            if units:
                print(f"  derived: {(rd + wr) / units:.1f} DRAM bytes per unit ({units:.0f} units per launch; read {rd / units:.1f}, write {wr / units:.1f})")
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
