#!/usr/bin/env python
"""Rank the ops of one profiled step (bench.py --dump-ops CSV) by the time they spend ABOVE their roofline.

  python profiles/analyze_ops.py profiles/r1y_ops.csv [peak_tflops] [peak_gbs]

Tensor ops ("tc:*") are compared with algorithmic FLOPs / sustained tensor peak, the others with algorithmic bytes / HBM peak
(defaults: MEASURED_PEAKS.json).  Per-op times come from CUDA events around every launch of one step, i.e. under the step's
power-capped clocks — the same clocks the sustained peak was measured at."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    peaks = {}
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        peaks = json.load(open(mp))
    tf = float(sys.argv[2]) if len(sys.argv) > 2 else peaks.get("bf16_tflops_sustained", 1400.0)
    gb = float(sys.argv[3]) if len(sys.argv) > 3 else peaks.get("hbm_gbs", 6550.0)
    rows = list(csv.DictReader(open(path)))
    fam = collections.OrderedDict()
    total = 0.0
    for r in rows:
        ms, gf, mbytes = float(r["ms"]), float(r["gflop"]), float(r["mbytes"])
        # GFLOP / (TFLOP/s) and MB / (GB/s) are both milliseconds
        ideal = gf / tf if r["kind"].startswith("tc:") and gf > 0 else mbytes / gb
        f = fam.setdefault(r["kind"], [0, 0.0, 0.0])
        f[0] += 1
        f[1] += ms
        f[2] += ideal
        total += ms
    print(f"# {os.path.basename(path)}: {len(rows)} ops, {total:.1f} ms; peaks {tf:.0f} TFLOP/s, {gb:.0f} GB/s")
    print(f"{'kind':26s} {'ops':>4s} {'ms':>8s} {'roofline ms':>12s} {'excess ms':>10s} {'of step':>8s} {'efficiency':>10s}")
    tot_ideal = 0.0
    for k, (n, ms, ideal) in sorted(fam.items(), key=lambda kv: -(kv[1][1] - kv[1][2])):
        tot_ideal += ideal
        print(f"{k:26s} {n:4d} {ms:8.2f} {ideal:12.2f} {ms - ideal:10.2f} {100 * (ms - ideal) / total:7.1f}% {100 * ideal / ms if ms else 0:9.1f}%")
    print(f"{'total':26s} {len(rows):4d} {total:8.2f} {tot_ideal:12.2f} {total - tot_ideal:10.2f} {100 * (total - tot_ideal) / total:7.1f}% {100 * tot_ideal / total:9.1f}%")


if __name__ == "__main__":
    main()
