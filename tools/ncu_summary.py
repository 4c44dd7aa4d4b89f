"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("b200lp::", "")
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':58s} {'n':>4s} {'mean us':>10s} {'min us':>10s} {'max us':>10s} {'share':>7s}")
for k, v in agg.items():
    print(f"{k:58s} {len(v):4d} {sum(v) / len(v):10.1f} {min(v):10.1f} {max(v):10.1f} {100 * sum(v) / tot:6.1f}%")
