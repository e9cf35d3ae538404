"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, opcode mix, stall reasons,
hottest SASS lines.  usage: python scripts/ncu_summary.py report.ncu-rep [cell_steps]"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    cell_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg.per_second",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
            "lts__t_bytes.sum", "lts__t_sectors_op_read.sum"]
    for k in keys:
        if k in m:
            print("%-75s %s %s" % (k, m[k], u.get(k, "")))
    print("\nstall reasons (warps per issue-active cycle):")
    for k in sorted(m):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
            print("  %-28s %s" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], m[k]))
    src = page(rep, "source")
    h, data = src[1], src[2:]
    ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    tot = sum(int(r[ia]) for r in data)
    tots = sum(int(r[isamp]) for r in data)
    print("\nwarp instructions executed: %d  (%d SASS lines)" % (tot, len(data)))
    if cell_steps:
        print("per cell-step: %.0f warp-instr = %.0f lane-instr" % (tot / cell_steps, 32 * tot / cell_steps))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        o = mm.group(2).split(".")[0] if mm else "?"
        op[o] += int(r[ia])
        ops[o] += int(r[isamp])
    print("\nopcode mix (% of executed instructions | % of stall samples):")
    for o, c in op.most_common(24):
        print("  %-10s %6.2f%%  %6.2f%%" % (o, 100.0 * c / tot, 100.0 * ops[o] / tots))
    print("\nhottest SASS lines by samples:")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:25]:
        print("  %6.2f%%  x%-10s %s" % (100.0 * int(r[isamp]) / tots, r[ia], r[isrc].strip()[:90]))


if __name__ == "__main__":
    main()
