"""Condenses one `ncu --set full` report of tools/prof_kernels.py into profiles/<out>.json / .txt: EVERY captured launch in order, matched with the
workload list the script wrote (gpurun_out/prof_kernels_workloads.json): duration, DRAM bytes against algorithmic bytes, tensor-pipe activity,
L1/shared and L2 throughput, registers, and the isolated (cold-cache, serialised) rate.
usage: python tools/ncu_summary2.py profiles/r02_ncu_kernels gpurun_out/r02_kernels.ncu-rep gpurun_out/prof_kernels_workloads.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_throughput_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__grid_size": "grid", "launch__block_size": "block", "launch__cluster_size": "cluster"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    out, rep, wl = sys.argv[1:4]
    work = json.load(open(wl))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res, lines, wi = {}, [], 0
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = name.replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
        d = {"kernel": short}
        for k, nk in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if nk.startswith("dram_r") or nk.startswith("dram_w"):
                    v *= UNIT.get(units[i], 1)
                if nk == "duration_us":
                    v = v / 1000.0 if units[i] == "ns" else (v * 1000.0 if units[i] == "ms" else v)
                d[nk] = v
        d["dram_bytes_per_launch"] = d.get("dram_read", 0) + d.get("dram_write", 0)
        # match with the next workload whose kernel base name agrees
        base = short.split("<")[0]
        label = None
        for j in range(wi, len(work)):
            if work[j]["kernel"].split("<")[0] == base:
                label, wi = work[j]["label"], j + 1
                d["family"] = work[j].get("family", base)
                d["algorithmic_flops_per_launch"] = work[j]["algorithmic_flops"]
                d["algorithmic_bytes_per_launch"] = work[j]["algorithmic_bytes"]
                break
        label = label or ("unmatched %d %s" % (len(res), short))
        t = d["duration_us"] * 1e-6
        if d.get("algorithmic_flops_per_launch"):
            d["tflops_isolated"] = d["algorithmic_flops_per_launch"] / t / 1e12
        if d.get("algorithmic_bytes_per_launch"):
            d["gbs_algorithmic_isolated"] = d["algorithmic_bytes_per_launch"] / t / 1e9
            d["dram_over_algorithmic"] = d["dram_bytes_per_launch"] / d["algorithmic_bytes_per_launch"]
        d["note"] = "ncu --set full --clock-control none, ONE cold launch (tools/prof_kernels.py)"
        d["workload"] = label
        res[label] = d
        lines.append("%-58s %-34s %8.1f us %7.0f TF/s  dram %7.1f MB (x%.2f alg)  tensor %5.1f%%  l1/smem %5.1f%%  L2 %5.1f%%  dram-tp %5.1f%%  regs %3d" % (
            label, short, d["duration_us"], d.get("tflops_isolated", 0), d["dram_bytes_per_launch"] / 1e6, d.get("dram_over_algorithmic", 0),
            d.get("tensor_pipe_active_pct", 0), d.get("l1tex_throughput_pct", 0), d.get("l2_throughput_pct", 0), d.get("dram_throughput_pct", 0),
            int(d.get("registers", 0))))
    with open(out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    with open(out + ".txt", "w") as f:
        f.write("# ncu --set full --clock-control none: one cold, serialised launch per row (tools/prof_kernels.py); rates are isolated-launch rates\n")
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
