"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of ONE denoise step (between two
consecutive cfg_euler_kernel launches), grouped by kernel.   python tools/launch_summary.py gpurun_out/step_launches.csv"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).strip()
    name = re.sub(r"^void\s+", "", name).replace("sa::", "")
    rows.append((name, v))
marks = [i for i, (n, _) in enumerate(rows) if "cfg_euler_kernel" in n]
if len(marks) >= 2:
    seg = rows[marks[-2] + 1:marks[-1] + 1]
    note = "one denoise step (between two cfg_euler_kernel launches)"
else:
    seg, note = rows, "whole capture (fewer than two cfg_euler_kernel launches found)"
tot, cnt = collections.Counter(), collections.Counter()
for n, v in seg:
    tot[n] += v
    cnt[n] += 1
T = sum(tot.values())
print(f"# {note}: {len(seg)} launches, {T:.1f} ms serialised / cold-cache")
for n, v in tot.most_common(24):
    print(f"{n[:78]:78s} n={cnt[n]:4d} {v:9.2f} ms {100 * v / T:5.1f}%")
