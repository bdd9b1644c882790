"""Hot source lines of an ncu --set full --import-source capture (read here, no GPU): warp-stall samples and executed
instructions aggregated per CUDA source line.

    python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [top]
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
fname = "?"
agg = defaultdict(lambda: [0, 0, defaultdict(int), ""])
i = 0
while i < len(lines):
    row = next(csv.reader([lines[i]]))
    if row and row[0] == "File Name":
        fname = row[1].split("/")[-1]
    if row and row[0] == "Line No":
        hdr = row
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        i += 1
        while i < len(lines):
            r = next(csv.reader([lines[i]]))
            if not r or r[0] in ("File Name", "Line No"):
                i -= 1
                break
            try:
                key = (fname, int(r[0]))
            except ValueError:
                i += 1
                continue
            a = agg[key]
            a[3] = r[1].strip()[:110]
            try:
                a[0] += int(r[si] or 0)
                a[1] += int(r[ii] or 0)
                for k, h in stall_cols:
                    a[2][h] += int(r[k] or 0)
            except (ValueError, IndexError):
                pass
            i += 1
    i += 1
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{h[6:]} {v}" for h, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100 * a[0] / tot_s:5.1f}% samples {100 * a[1] / tot_i:5.1f}% instr  {f}:{ln:<5d} {a[3]}\n        [{st}]")
