"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel and, for the BCR kernels, per level."""
import collections
import csv
import sys

path = sys.argv[1]
allrows = list(csv.reader(open(path)))
hdr = [r for r in allrows if len(r) > 10 and r[0] == "ID"][0]
rows = [r for r in allrows if len(r) > 10 and r[0].isdigit()]
gi = hdr.index("Grid Size")
d = collections.defaultdict(lambda: [0, 0.0])
lv = collections.defaultdict(list)
for r in rows:
    name = r[4].split("(")[0].replace("acino::", "")
    v = float(r[-1].replace(",", "")) / 1e3
    d[name][0] += 1
    d[name][1] += v
    if "bcr" in name or "fte_jac" in name:
        lv[(name, int(r[gi].strip("()").split(",")[0]))].append(v)
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{k[:56]:56s} n={v[0]:5d} total={v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  avg={v[1] / v[0]:8.1f} us")
for k, v in sorted(lv.items()):
    print(f"  {k[0]:22s} grid {k[1]:5d}: n={len(v):3d} avg {sum(v) / len(v):8.1f} us  min {min(v):8.1f}")
