"""Per-launch table from an `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,
dram__bytes_read.sum,dram__bytes_write.sum --csv` log: python tools/ncu_launch_table.py in.csv out.txt "title" [top]"""
import collections
import csv
import sys

src, dst, title = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 70
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
L = collections.OrderedDict()
for r in data:
    d = L.setdefault(r[ii], {"name": r[ki]})
    v, u = float(r[vi].replace(",", "")), r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    if r[mi].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r[mi]] = v
out = []
for i, d in enumerate(L.values()):
    nm = d["name"].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
    out.append((i, nm, d.get("gpu__time_duration.sum", 0.0), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
                (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e6))
tot = sum(r[2] for r in out)
with open(dst, "w") as f:
    f.write("# %s\n# every launch in order under ncu --clock-control none (cold, serialised): duration, tensor-pipe activity, DRAM bytes\n"
            "# total %.1f us over %d launches\n" % (title, tot, len(out)))
    for r in sorted(out, key=lambda r: -r[2])[:top]:
        f.write("#%3d %-46s %9.1f us %5.1f%%  tensor pipe %5.1f%%  dram %8.1f MB  %6.0f GB/s\n" % (
            r[0], r[1][:46], r[2], 100 * r[2] / tot, r[3], r[4], r[4] / max(r[2], 1e-9) * 1000.0))
print(open(dst).read()[:1500])
