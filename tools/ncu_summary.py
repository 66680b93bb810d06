"""Summarise an .ncu-rep (ncu --set full) into a small table: python tools/ncu_summary.py report.ncu-rep [out.md]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1 %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "lim smem"),
    ("launch__occupancy_limit_registers", "lim regs"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st long_sb"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "st short_sb"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "st barrier"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "st wait"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "st math"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "st lg"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "st mio"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "st membar"),
    ("smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "st branch"),
    ("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "st noinst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        grid, block = r[hdr.index("Grid Size")], r[hdr.index("Block Size")]
        lines.append("### %s  grid %s block %s" % (name, grid, block))
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                lines.append("| %s (`%s`) | %s | %s |" % (label, key, r[i], units[i]))
        # warp-stall breakdown: share of the sampled warp states (smsp__pcsamp_warps_issue_stalled_*), largest first
        st = []
        for i, key in enumerate(hdr):
            if key.startswith("smsp__pcsamp_warps_issue_stalled_") and not key.endswith("_not_issued"):
                try:
                    st.append((float(r[i].replace(",", "")), key[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st)
        if tot > 0:
            lines.append("")
            lines.append("warp states (pc sampling): " + ", ".join("%s %.1f %%" % (k, 100.0 * v / tot) for v, k in sorted(st, reverse=True)[:8]))
        lines.append("")
    out = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(out + "\n")
    print(out)


if __name__ == "__main__":
    main()
