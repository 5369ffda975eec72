"""Turn gpurun_out ncu artefacts into small, committed summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/NAME_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof_X.ncu-rep|prof_X_raw.csv profiles/NAME_full.csv [profiles/traffic.json]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    data = [(r[i_name], float(r[i_val].replace(",", ""))) for r in rows[1:] if r[i_id].isdigit()]
    marks = [i for i, (n, _) in enumerate(data) if "instnorm" in n]
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none launch list (cold-cache, serialised):",
           "# compare SHARES of the step, not absolutes.  source: %s, %d launches, steps start at the instnorm kernel" % (src, len(data))]
    for si in range(len(marks) - 1):
        step = data[marks[si]:marks[si + 1]]
        tot = sum(v for _, v in step)
        agg = collections.OrderedDict()
        for n, v in step:
            key = re.sub(r"\(.*", "", n)[:90]
            c, t = agg.get(key, (0, 0.0))
            agg[key] = (c + 1, t + v)
        out.append("\n== step %d: %d launches, %.3f ms total" % (si, len(step), tot / 1e6))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
            out.append("%6.2f%% %10.3f ms x%3d  %s" % (t / tot * 100, t / 1e6, c, k))
    open(dst, "w").write("\n".join(out) + "\n")


def full(src, dst, traffic=None):
    # src: an .ncu-rep, or the `ncu -i X.ncu-rep --page raw --csv` export made on the GPU box (reports of a whole step
    # exceed what a gpurun call may bring back; the export is a few hundred KB)
    raw = open(src).read() if src.endswith(".csv") else \
        subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines()) if r]
    first = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    rows = rows[first:]
    hdr, units = rows[0], rows[1]
    cols = [h for h in KEEP if h in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[hdr.index(c)] for c in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index(c)] for c in cols])
    if traffic:
        agg = {}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            fams = ["shapelet_fwd"] if "shapelet_fwd" in name else \
                ["shapelet_bwd", "bwd.contraction"] if ("shapelet_bwd_kernel" in name or "shapelet_bwd_tc_kernel" in name) else \
                ["bwd.pool_bwd"] if "pool_bwd" in name else ["bwd.tie_check"] if "tie_check" in name else \
                ["instnorm"] if "instnorm" in name else ["window_prefix"] if "prefix" in name else \
                ["window_stats"] if "window_stats" in name else []
            if not fams:
                continue

            def tobytes(col):
                v, u = float(r[hdr.index(col)].replace(",", "")), units[hdr.index(col)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            for fam in fams:
                n, t = agg.get(fam, (0, 0.0))
                agg[fam] = (n + 1, t + tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
        json.dump({k: t / n for k, (n, t) in agg.items()}, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
