#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the text summaries committed under profiles/.

  python profiles/summarize.py <tag>      # reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_*_<tag>.ncu-rep
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def launches(tag, out):
    path = os.path.join(GO, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("sdm::", "")
        t = float(r[mv].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
        n += 1
    out.write(f"# ncu launch list `{os.path.basename(path)}`: {n} launches, {total / 1e3:.2f} ms summed kernel time\n")
    out.write("# (ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes)\n")
    out.write(f"{'kernel':58s} {'launches':>8s} {'ms':>10s} {'share':>7s}\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{k[:58]:58s} {c:8d} {t / 1e3:10.3f} {100 * t / total:6.1f}%\n")


def reports(tag, out):
    for rep in sorted(glob.glob(os.path.join(GO, f"prof_*_{tag}.ncu-rep"))):
        csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(csvtxt.splitlines()))
        if not rows:
            continue
        hdr, units = rows[0], rows[1]
        out.write(f"\n# ncu --set full capture `{os.path.basename(rep)}`\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            out.write(f"kernel: {d.get('Kernel Name', '?')}  grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}\n")
            for m in METRICS:
                if m in d:
                    u = units[hdr.index(m)]
                    out.write(f"    {m:72s} {d[m]:>14s} {u}\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    dst = os.path.join(ROOT, "profiles", f"{tag}_summary.txt")
    with open(dst, "w") as f:
        launches(tag, f)
        reports(tag, f)
    print(open(dst).read())
