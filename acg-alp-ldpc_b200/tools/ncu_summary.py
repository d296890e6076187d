"""Condenses an Nsight Compute report (.ncu-rep) into the text summary that is
committed under profiles/ (the reports themselves stay in gpurun_out/).

    python acg-alp-ldpc_b200/tools/ncu_summary.py gpurun_out/prof.ncu-rep "units per launch" > profiles/rNN_x.txt

`units` (optional) = algorithmic work units of the profiled launch (edge-iterations
for BP, block-iterations for QP-ADMM) to express instruction counts per unit.
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
d = dict(zip(raw[0], zip(raw[1], raw[2])))
print("# ncu summary of", rep)
keys = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
]
for k in keys:
    if k in d:
        print("%-75s %s %s" % (k, d[k][1], d[k][0]))
print("\n# warp stall reasons (warp-cycles per issued instruction)")
for k in sorted(d):
    if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
        try:
            v = float(d[k][1])
        except ValueError:
            continue
        if v >= 0.05:
            print("%-75s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
src = page("source")
hdr, data = src[1], src[2:]
iA, iT, iE = hdr.index("Source"), hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed")
ops = collections.Counter()
for r in data:
    s = r[iA].split()
    o = (s[0] if not s[0].startswith("@") else s[1]).split(".")[0]
    ops[o] += int(r[iT])
tot = sum(ops.values())
warp_inst = sum(int(r[iE]) for r in data)
print("\n# instruction mix (thread instructions; %d SASS lines; lane utilisation %.3f)" % (len(data), tot / (32.0 * warp_inst)))
if units:
    print("# per algorithmic unit (%g units in this launch): %.1f thread instructions" % (units, tot / units))
for o, c in ops.most_common(24):
    print("%-10s %6.2f%%%s" % (o, 100.0 * c / tot, ("   %7.2f / unit" % (c / units)) if units else ""))
