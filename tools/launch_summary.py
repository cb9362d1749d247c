"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: kernels after the LAST launch whose name
starts with argv[2], grouped by name.     python tools/launch_summary.py launches.csv k_ipa_begin"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
names = [r[4].replace("vk::", "").replace("void ", "") for r in rows]
start = max(i for i, n in enumerate(names) if n.startswith(sys.argv[2]))
agg = collections.OrderedDict()
tot = 0.0
for r, nm in zip(rows[start:], names[start:]):
    n = nm.split("(")[0]
    d = float(r[-1].replace(",", "")) / 1e3
    tot += d
    c, s = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, s + d)
print(f"{len(rows) - start} launches, {tot:.1f} us of kernel time (ncu: cold caches, serialised)")
for n, (c, s) in agg.items():
    print(f"  {n:50s} x{c:3d} {s:9.1f} us ({s / c:7.1f} each)")
