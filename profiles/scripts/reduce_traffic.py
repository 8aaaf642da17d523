"""ncu --csv launch list (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> per-kernel-family totals.
The script that produced the list runs two identical forwards (warm-up + measured): every figure is halved to one forward."""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    rows.append(r)
unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
fam = {}
for r in rows:
    name = r["Kernel Name"]
    if "conv_swap" in name or "(bool)1>" in name.split("conv_gemm_kernel")[-1][-12:]:
        k = "conv3x3"  # conv_swap*_kernel and the HALO instantiations of conv_gemm_kernel (last template argument true) are 3x3 stride-1 convs
    else:
        k = "other_gemm"
    e = fam.setdefault(k, {"dram_bytes": 0.0, "ms": 0.0, "launches": 0})
    v = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r["Metric Unit"], 1.0)
    if r["Metric Name"].startswith("dram__bytes"):
        e["dram_bytes"] += v
    elif r["Metric Name"].startswith("gpu__time_duration"):
        e["ms"] += v
        e["launches"] += 1
forwards = 2
out = {k: {"dram_bytes": v["dram_bytes"] / forwards, "ms_cold_serialised": v["ms"] / forwards, "launches": v["launches"] // forwards} for k, v in fam.items()}
out["note"] = "ncu --clock-control none, per-launch replay (cold caches, serialised); bs=8 1024^2, one forward; conv3x3 = conv_swap*_kernel + HALO instantiations of conv_gemm_kernel"
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
