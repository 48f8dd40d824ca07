"""Summarise ncu outputs under gpurun_out/ into small text/JSON files under profiles/ (run in the build container)."""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = OrderedDict()
        d["kernel"] = vals[hdr.index("Kernel Name")]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = "%s %s" % (vals[i], units[i])
        res.append(d)
    return res


def launches(path):
    agg = defaultdict(list)
    order = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        if name not in agg:
            order.append(name)
        agg[name].append(v_us)
    return order, agg


if __name__ == "__main__":
    tag = sys.argv[1]
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    go = os.path.join(ROOT, "gpurun_out")
    lines = []
    lp = os.path.join(go, tag + "_launches.csv")
    if not os.path.exists(lp):
        lp = os.path.join(go, "launches.csv")
    if os.path.exists(lp):
        order, agg = launches(lp)
        tot = sum(sum(v) / len(v) for v in agg.values())
        lines.append("# launch list (ncu --metrics gpu__time_duration.sum --clock-control none): mean us per launch, share of one step")
        for n in order:
            v = agg[n]
            lines.append("%-36s launches=%3d mean=%10.1f us  share=%5.1f %%" % (n, len(v), sum(v) / len(v), 100 * sum(v) / len(v) / tot))
        lines.append("")
    for rep in sys.argv[2:]:
        p = os.path.join(go, rep + ".ncu-rep")
        if not os.path.exists(p):
            continue
        for d in raw(p):
            lines.append("# ncu --set full: %s  (%s)" % (rep, d["kernel"][:110]))
            for k, v in d.items():
                if k != "kernel":
                    lines.append("  %-86s %s" % (k, v))
            lines.append("")
    with open(os.path.join(out_dir, tag + "_ncu_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
