#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch, `ncu --set full --import-source on`) as markdown:
headline counters, stall reasons and the source lines that execute the most instructions.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.md ["title"]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_average_branch_targets_threads_uniform.pct"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = ["# " + title, "", "Source: `%s` (`ncu --set full --clock-control none --import-source on`)." % rep, ""]
    for k, vals in enumerate(raw[2:]):
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines += ["## launch %d: `%s`" % (k, d.get("Kernel Name", "?")), "", "| counter | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in d:
                lines.append("| %s | %s | %s |" % (w, d[w], u[w]))
        lines.append("")
        # warps per issued instruction waiting for each reason (the raw page's ratios; the per-line stall columns of the
        # source page are not aligned between its cuda and sass rows)
        pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        st = sorted(((float(v), k[len(pre):-len(suf)]) for k, v in d.items() if k.startswith(pre) and k.endswith(suf) and v), reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        lines += ["| stall reason | warps per issue | share |", "|---|---|---|"]
        lines += ["| %s | %.2f | %.1f%% |" % (k, v, 100 * v / tot) for v, k in st[:10]]
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, h, recs, stall = None, None, [], {}
    num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            h = r
        elif h and r[0].strip().isdigit():
            d = dict(zip(h, r))
            recs.append((num(d["Instructions Executed"]), num(d["# Samples"]), cur, int(r[0]), r[1].strip()[:100]))
            for name in h:
                if name.startswith("stall_") and "Not Issued" not in name:
                    stall[name] = stall.get(name, 0) + num(d[name])
    ti = max(1, sum(x[0] for x in recs))
    ts = max(1, sum(x[1] for x in recs))
    st = max(1, sum(stall.values()))
    lines += ["", "## source lines by warp instructions executed (total %d)" % ti, "",
              "| inst | samples | file:line | source |", "|---|---|---|---|"]
    for x in sorted(recs, key=lambda x: -x[0])[:30]:
        lines.append("| %.2f%% | %.2f%% | %s:%d | `%s` |" % (100.0 * x[0] / ti, 100.0 * x[1] / ts, x[2], x[3],
                                                              x[4].replace("|", "\\|")))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
