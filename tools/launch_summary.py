"""Developer tool: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
  python tools/launch_summary.py gpurun_out/launches.csv profiles/rX_launch_list_summary.csv "header comment"
"""
import csv, sys
from collections import defaultdict

src, dst, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1.0)
    agg[r[ki][:90]][0] += 1
    agg[r[ki][:90]][1] += v
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as fh:
    for line in note.split("\\n"):
        fh.write(f"# {line}\n")
    fh.write("kernel,launches,total_us,avg_us,share_pct\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f'"{k}",{n},{t:.1f},{t / n:.2f},{100 * t / tot:.1f}\n')
