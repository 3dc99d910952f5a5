"""Condenses `ncu --set full` reports into profiles/<out>.json / .txt: per kernel name the first captured launch's duration, DRAM bytes,
tensor-pipe activity, DRAM throughput and registers.  python tools/ncu_summary.py out_prefix label=report.ncu-rep ..."""
import csv
import io
import json
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__grid_size": "grid", "launch__block_size": "block"}


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    res, lines = {}, []
    for item in reps:
        label, path = item.rsplit("=", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = name.replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
            if short in seen:
                continue
            seen.add(short)
            d = {"report": path.split("/")[-1], "workload": label}
            for k, nk in WANT.items():
                if k in hdr:
                    i = hdr.index(k)
                    d[nk] = to_bytes(r[i], units[i]) if nk.startswith("dram_r") or nk.startswith("dram_w") else float(r[i].replace(",", ""))
                    if nk == "duration_us":
                        d[nk] = d[nk] / 1000.0 if units[i] == "ns" else (d[nk] * 1000.0 if units[i] == "ms" else d[nk])
            d["dram_bytes_per_launch"] = d.get("dram_read", 0) + d.get("dram_write", 0)
            res.setdefault(short, d)
            lines.append("%-44s %-34s %9.1f us  dram %7.1f MB rd %7.1f MB wr  tensor %5.1f%%  dram-tp %5.1f%%  regs %3d" % (
                short, label, d["duration_us"], d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6,
                d.get("tensor_pipe_active_pct", 0), d.get("dram_throughput_pct", 0), int(d.get("registers", 0))))
    with open(out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    with open(out + ".txt", "w") as f:
        f.write("# ncu --set full --clock-control none (one launch per kernel; cold L2, serialised): duration, DRAM traffic, tensor pipe\n")
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
