#!/usr/bin/env python
"""Compact summary of an .ncu-rep (ncu --set full) for profiles/: one block per kernel launch."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % (inst issue)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU/conversions) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA(FP32/IMAD) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "DFMA thread-inst / cycle (all SMs)"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "DMUL thread-inst / cycle (all SMs)"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "DADD thread-inst / cycle (all SMs)"),
    ("sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained", "FP64 peak thread-inst / cycle (all SMs)"),
    ("smsp__cycles_elapsed.avg", "cycles elapsed"),
    ("smsp__cycles_active.avg", "SMSP cycles active"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global/L1)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (MUFU/smem)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/smem data pipe % of peak"),
    ("idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed", "constant cache (IDC) % of peak"),
    ("sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed", "uniform pipe %"),
]


def main(path, points=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"kernel: {name[:150]}")
        vals = {}
        for k, label in KEYS:
            if k in idx:
                vals[k] = r[idx[k]]
                print(f"  {label:42s} {r[idx[k]]:>16s} {units[idx[k]]}")
        try:
            f = lambda k: float(vals[k].replace(",", ""))
            cyc = f("smsp__cycles_elapsed.avg")
            dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
            dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
            dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
            tot = f("smsp__inst_executed.sum") * f("smsp__thread_inst_executed_per_inst_executed.ratio")
            print(f"  {'FP64 arithmetic thread-inst (DFMA+DMUL+DADD)':42s} {dfma + dmul + dadd:16.4g}")
            print(f"  {'FP64 flops (DFMA = 2)':42s} {2 * dfma + dmul + dadd:16.4g}")
            if points:
                print(f"  per grid point ({points} points): DFMA {dfma/points:.1f} DMUL {dmul/points:.1f} DADD {dadd/points:.1f} "
                      f"-> FP64 arith inst {(dfma+dmul+dadd)/points:.1f}, flops {(2*dfma+dmul+dadd)/points:.1f}, "
                      f"all thread-inst {tot/points:.1f}, dram bytes {(f('dram__bytes_read.sum')+f('dram__bytes_write.sum'))*1e6/points:.1f}")
        except Exception as e:  # noqa
            print("  (derived figures unavailable:", e, ")")
        print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
