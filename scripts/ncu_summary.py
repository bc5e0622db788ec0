"""Condenses an .ncu-rep (ncu --set full --import-source on) into the text summary committed under
profiles/: headline counters per captured launch, SASS opcode mix with stall samples, and the
hottest source lines.  Runs on the CPU box (ncu -i).

    python scripts/ncu_summary.py gpurun_out/r01a/prof_tau.ncu-rep > profiles/r01a_k_tau.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, top=25):
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    print("# ncu summary of %s" % rep)
    for n, row in enumerate(raw[2:]):
        name = row[hdr.index("Kernel Name")]
        print("\n## launch %d: %s" % (n, name[:100]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-88s %s %s" % (k, row[i], units[i]))
    src = ncu_csv(rep, "source")
    heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    for n, h in enumerate(heads):
        cols = src[h]
        end = heads[n + 1] - 1 if n + 1 < len(heads) else len(src)
        body = [r for r in src[h + 1:end] if len(r) == len(cols)]
        c = {k: cols.index(k) for k in ("Source", "# Samples", "Instructions Executed", "stall_wait", "stall_long_sb",
                                        "stall_short_sb", "stall_math", "stall_branch_resolving", "stall_no_inst")}
        tot_s = sum(int(r[c["# Samples"]] or 0) for r in body) or 1
        tot_e = sum(int(r[c["Instructions Executed"]] or 0) for r in body) or 1
        ex, sm = collections.Counter(), collections.Counter()
        stall = collections.Counter()
        for r in body:
            tok = r[c["Source"]].split()
            op = (tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "")).split(".")[0]
            ex[op] += int(r[c["Instructions Executed"]] or 0)
            sm[op] += int(r[c["# Samples"]] or 0)
            for k in ("stall_wait", "stall_long_sb", "stall_short_sb", "stall_math", "stall_branch_resolving", "stall_no_inst"):
                stall[k] += int(r[c[k]] or 0)
        print("\n## launch %d SASS: %d instructions, %d warp-instructions executed, %d stall samples" % (n, len(body), tot_e, tot_s))
        print("stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in stall.most_common()))
        print("%-12s %10s %10s" % ("opcode", "executed%", "samples%"))
        for op, v in ex.most_common(top):
            print("%-12s %9.2f%% %9.2f%%" % (op, 100.0 * v / tot_e, 100.0 * sm[op] / tot_s))


if __name__ == "__main__":
    main(sys.argv[1])
