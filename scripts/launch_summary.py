"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py --steps 1 --warmup 1`:
launches between the last two vit_patchify_kernel launches = one full step; shares per kernel (per-launch times under
ncu are cold-cache and serialised: compare SHARES, not absolutes).   python scripts/launch_summary.py launches.csv > out.json"""
import collections
import csv
import json
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))  # us
marks = [i for i, (k, _) in enumerate(rows) if "vit_patchify_kernel" in k]
assert len(marks) >= 2, "need two steps in the capture"
step = rows[marks[-2]:marks[-1]]
tot = sum(t for _, t in step)
agg = collections.OrderedDict()
for k, t in step:
    name = re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
kern = sorted(({"kernel": k, "launches": n, "total_us": round(t, 1), "share": round(t / tot, 4)} for k, (n, t) in agg.items()),
              key=lambda d: -d["total_us"])
print(json.dumps({"command": "ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 900 --csv python bench.py "
                             "--steps 1 --warmup 1 --cpu-clips 2 --no-extras",
                  "note": "launches between two consecutive vit_patchify_kernel launches = one full step; per-launch times are "
                          "cold-cache and serialised: compare SHARES",
                  "launches_in_step": len(step), "total_us": round(tot, 1), "kernels": kern}, indent=1))
