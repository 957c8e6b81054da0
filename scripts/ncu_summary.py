#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) into the numbers we track.  Runs on the CPU box.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--md profiles/name.md]
"""
import collections, csv, io, subprocess, sys

def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
    lines = []
    for k in keys:
        if k in m:
            lines.append("%-75s %s %s" % (k, m[k], u.get(k, "")))
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            if float(m[h] or 0) > 0.05:
                lines.append("%-75s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall/issue: ").replace("_per_issue_active.ratio", ""), m[h]))
    src = page(rep, "source")
    if len(src) > 2:
        h2 = src[1]; ix = {n: i for i, n in enumerate(h2)}
        tot = collections.Counter(); cnt = collections.Counter(); wf = collections.Counter(); wfi = collections.Counter()
        total = 0
        for r in src[2:]:
            s = r[ix["Source"]].split()
            if not s: continue
            op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
            n = int(r[ix["# Samples"]] or 0); ex = int(r[ix["Instructions Executed"]] or 0)
            tot[op] += n; cnt[op] += ex; total += n
            w = r[ix["L1 Wavefronts Shared"]]
            if w: wf[op] += int(w); wfi[op] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        lines.append("")
        lines.append("opcode        samples    %   executed(warp-inst)  smem wavefronts (ideal)")
        for op, s in tot.most_common(16):
            lines.append("%-12s %8d %5.1f %14d %14d (%d)" % (op, s, 100.0 * s / max(total, 1), cnt[op], wf[op], wfi[op]))
        lines.append("total warp-instructions executed: %d" % sum(cnt.values()))
    text = "\n".join(lines)
    print(text)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write("```\n" + text + "\n```\n")

if __name__ == "__main__":
    main()
