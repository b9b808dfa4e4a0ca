"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0][:70]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print("%-72s %5s %12s %12s %6s" % ("kernel", "n", "avg_us", "sum_us", "share"))
for k, v in agg.items():
    print("%-72s %5d %12.1f %12.1f %5.1f%%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
