"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/summarize_launches.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr, data = None, []
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    data.append(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in data:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    name = r[ki].replace("void <unnamed>::", "").replace("<unnamed>::", "")[:64]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print("launches %d total_us %.1f" % (len(data), tot))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-64s n=%4d %10.1f us %5.1f%%" % (k, n, t, 100 * t / tot))
