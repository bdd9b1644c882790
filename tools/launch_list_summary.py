"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list (read here, no GPU):

    python tools/launch_list_summary.py gpurun_out/x.csv "command line that was profiled" > profiles/r02_launch_list_summary.txt
"""
import collections
import csv
import re
import sys

path, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.reader(lines)
hdr = None
for r in rd:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    if len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.sub(r"dig3d::", "", name)
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9}.get(unit, 1.0)
    tot[name][0] += 1
    tot[name][1] += v
grand = sum(v[1] for v in tot.values()) or 1.0
print(what)
print("(cold-cache, serialised under ncu: compare SHARES, not absolute times)")
print(f"{'kernel':86s} {'launches':>8s} {'total ns':>14s} {'share':>7s}")
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:86]:86s} {n:8d} {ns:14.1f} {100 * ns / grand:6.1f}%")
print(f"{'TOTAL':86s} {sum(v[0] for v in tot.values()):8d} {grand:14.1f}")
ours = sum(ns for name, (n, ns) in tot.items() if not name.startswith("at::") and "nccl" not in name.lower() and "cub::" not in name)
print(f"share of dig_b200 kernels (libdig3d.so): {100 * ours / grand:.1f}%   ATen / library kernels: {100 * (grand - ours) / grand:.1f}%")
