"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list per kernel.
    python tools/ncu_summary.py gpurun_out/r02_launches.csv > profiles/r02_launch_summary.md"""
import collections
import csv
import re
import sys

rows = collections.OrderedDict()
with open(sys.argv[1], newline="") as f:
    lines = f.readlines()
start = next(i for i, l in enumerate(lines) if l.lstrip('"').startswith("ID"))     # ncu prints ==PROF== lines before the CSV
lines = lines[start:]
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |tb::|\(anonymous namespace\)::|<unnamed>::", "", name)
    if "at::native" in name or "vectorized_elementwise" in name or "distribution" in name or "philox" in name.lower():
        continue                                     # torch's random-weight generation (profiling tool only)
    key = (name, r["Grid Size"])
    d = rows.setdefault(key, {"n": set(), "t": 0.0, "rd": 0.0, "wr": 0.0})
    d["n"].add(r["ID"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["t"] += v / 1e3 if unit == "ns" else v * (1e3 if unit == "ms" else 1.0) if unit in ("ms", "us", "usecond") else v / 1e3
    else:
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        d["rd" if "read" in r["Metric Name"] else "wr"] += v * scale
total = sum(d["t"] for d in rows.values())
print("| kernel | grid | launches | avg us | share % | DRAM read MB/launch | DRAM write MB/launch |")
print("|---|---|---|---|---|---|---|")
for (name, grid), d in sorted(rows.items(), key=lambda kv: -kv[1]["t"]):
    n = len(d["n"])
    print(f"| {name} | {grid} | {n} | {d['t'] / n:.2f} | {100 * d['t'] / total:.1f} | {d['rd'] / n:.2f} | {d['wr'] / n:.2f} |")
