#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: train images/sec of one full G+D step at 256 x 256.

  python bench.py --gpus N --steps K --warmup W            our arm (libshmgan kernels on B200s)
  python bench.py --impl reference --gpus N ...            the CPU arm: the oracle port of the reference's step on host cores

One "image" = one polarimetric sample = the 5-tuple (I0, I45, I90, I135, ED) consumed by one train_step slot
(6 generator passes, 12 discriminator passes, 1 SpecSeg pass, both backward sweeps, clip + Adam).
A "step" = one train_step call on a batch of `--batch` samples per GPU (weak scaling: per-GPU work fixed as N grows).
For N > 1 launch with torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_images_per_sec_GD_step_256"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-extras", action="store_true", help="skip the roofline / cpu_baseline / parity-mode / inference legs")
    ap.add_argument("--ref-size", type=int, default=256, help="image side of the CPU arm's bounded sample (rate rescaled to 256 if smaller)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def make_config(B, S, dtype, world):
    """The workload both arms are measured on (BASELINE.json configs[1] shape; configs[2] at world > 1)."""
    return {"workload": "batch %d/GPU @%dx%d, %s mode: one G+D train step (6 G + 12 D + 1 SpecSeg passes, both backward sweeps, "
                        "clip+Adam) on 4 polarimetric images + pseudo-diffuse, live mask; configs[1] shape%s%s"
                        % (B, S, S, "bf16 tcgen05" if dtype == "bf16" else "fp32 parity",
                           " (configs[1]'s literal fp32 parity mode is timed beside it: parity_mode_fp32)" if dtype == "bf16" else "",
                           "" if world == 1 else "; configs[2] data-parallel, global batch %d" % (B * world)),
            "batch_per_gpu": B, "global_batch": B * world, "image_size": S, "parallelism": "dp%d" % world,
            "l2": "inputs+activations of one step >> 126 MB L2 (no flush needed)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's train step on the host cores
# ------------------------------------------------------------------------------------------------------------------------
def cpu_train_step_fn(size, fs=64):
    """Returns (fn, cores): fn() runs ONE sample (B=1) of the reference train step -- forward of all 6 G / 12 D / 1 SpecSeg
    passes, both gradient sweeps, clip + Keras Adam -- in float32 with every host thread torch can use."""
    import torch
    import oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    f = torch.float32
    Gp = O.init_params(O.generator_param_specs(fs, True), 42, f)
    Dp = O.init_params(O.discriminator_param_specs(size, fs, True), 43, f)
    Sp = O.init_params(O.specseg_param_specs(), 44, f)
    g = torch.Generator().manual_seed(0)
    pol = [torch.rand((1, size, size, 3), generator=g) for _ in range(4)]
    origs = pol + [O.pseudo_diffuse_min4(*pol)]
    state = {"G": Gp, "D": Dp, "step": 0,
             "mG": {k: torch.zeros_like(v) for k, v in Gp.items()}, "vG": {k: torch.zeros_like(v) for k, v in Gp.items()},
             "mD": {k: torch.zeros_like(v) for k, v in Dp.items()}, "vD": {k: torch.zeros_like(v) for k, v in Dp.items()}}

    def fn():
        with torch.no_grad():
            Y90 = O.per_image_standardization(O.rgb_to_yuv(origs[2]), True)[0][..., 0:1]
            mask = O.specseg_forward(Sp, Y90)
        noise = [torch.randn((1, size, size, 3), generator=g) * 0.1 for _ in range(2)]
        keep = [(torch.rand((1, size // 32, size // 32, fs * 16), generator=g) < 0.8).float() for _ in range(2)]
        L, gG, gD = O.train_step_grads(state["G"], state["D"], origs, mask, [True, False, True, False, False], 0.9, noise, keep)
        with torch.no_grad():
            for key, grads, m, v in (("D", gD, "mD", "vD"), ("G", gG, "mG", "vG")):
                P = {k: state[key][k] for k in grads}
                P, mm, vv = O.keras_adam_update(P, grads, {k: state[m][k] for k in grads}, {k: state[v][k] for k in grads}, state["step"])
                state[key].update(P); state[m].update(mm); state[v].update(vv)
        state["step"] += 1
        return float(L["total_Generator_loss"])
    return fn, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, cores = cpu_train_step_fn(a.ref_size)
    for _ in range(min(a.warmup, 1)):                      # one warm-up sample is enough on the CPU (no clocks / caches to settle)
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    val = a.steps / dt * (a.ref_size / 256.0) ** 2          # FLOPs of the fully-convolutional step scale with the pixel count
    sample = ("%d train steps of ONE sample (B=1) at %dx%d, float32, PyTorch-CPU oracle port (TensorFlow is not installable here), "
              "%d host threads%s" % (a.steps, a.ref_size, a.ref_size, cores,
                                     "; reported rate = measured rate x (size/256)^2 (work scales with the pixel count)" if a.ref_size != 256 else ""))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same workload description as our arm; the CPU steps a BOUNDED SAMPLE of it (one of the batch's samples per step: see cpu_baseline.sample)
            "config": make_config(a.batch, a.size, a.dtype, max(1, a.gpus)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from shmgan_b200 import _lib, model as M, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == a.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % a.gpus
    B, S = a.batch, a.size
    hbm, tf_burst, tf_sus, peak_src = peaks()

    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype=a.dtype, allow_random_specseg=True).build()
    if world > 1:
        net.enable_data_parallel(bucket_mb=float(os.environ.get("SHM_DP_BUCKET_MB", "25")))
        net.dp_overlap = os.environ.get("SHM_DP_OVERLAP", "1") != "0"
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    pol = [torch.rand((B, S, S, 3), generator=g, device="cuda") for _ in range(4)]
    dev_in = pol + [net.calculate_estimate_diffuse(*pol)]
    host_in = [t.cpu().pin_memory() for t in dev_in]
    stage = [torch.empty_like(t) for t in dev_in]

    def trace(msg):
        if os.environ.get("BENCH_TRACE"):
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rank_ms = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            per_rank = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(per_rank, ms)
            rank_ms.clear(); rank_ms.extend(float(t) / steps for t in per_rank)      # every rank's own device time per step (skew diagnosis)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident():
        net.train_step(*dev_in)

    loss_box = [0.0]

    # e2e: every step's five input tensors come from pinned host memory.  The copies are double-buffered on a side stream (what a
    # prefetching loader does): the H2D transfer of step i+1 runs while step i computes; step i waits for ITS copy before it starts.
    copy_stream = torch.cuda.Stream()
    stages = [stage, [torch.empty_like(t) for t in dev_in]]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_i = [0]

    def issue_copy(buf):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[buf])              # the step that last read this buffer has finished
            for s, h in zip(stages[buf], host_in):
                s.copy_(h, non_blocking=True)               # pinned host -> device, one full input set per step
            ready[buf].record(copy_stream)

    def step_e2e():
        buf = e2e_i[0] & 1
        torch.cuda.current_stream().wait_event(ready[buf])
        issue_copy(buf ^ 1)                                 # next step's inputs, overlapped with this step's compute
        net.train_step(*stages[buf])
        freed[buf].record()
        loss_box[0] = net.total_Generator_loss              # already read back from the device by train_step (loss table D2H)
        e2e_i[0] += 1

    # CUDA-graph replay of the step (net.cuda_graph; SHM_CUDA_GRAPH=0 keeps the eager launches): one graph per drop-bit pattern, so the
    # warm-up walks all 32 patterns once -- every TIMED step, with its own random pattern, is then a replay.
    use_graph = os.environ.get("SHM_CUDA_GRAPH", "1") != "0"
    net.cuda_graph = use_graph
    for _ in range(max(a.warmup, 3)):
        step_resident()
    if use_graph:
        for pat in range(32):
            net.drop_bits = [bool((pat >> j) & 1) for j in range(5)]
            step_resident()
        net.drop_bits = None
        torch.cuda.synchronize()
        use_graph = net.cuda_graph                          # False if the capture failed and the model fell back to eager launches
        trace("graphs captured: %d" % len(net._graphs))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launches()
    trace("warm-up done")
    ms = timed(step_resident, a.steps)
    trace("timed steps done")
    launches = (_lib.launches() - l0) // a.steps
    clk = clocks.stop() if rank == 0 else None
    value = world * B * a.steps / (ms * 1e-3)
    rank_ms_train = list(rank_ms)

    freed[0].record(); freed[1].record()
    issue_copy(0)
    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    trace("e2e done")
    e2e = {"value": world * B * a.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_in), "d2h_bytes_per_step": net.table.buf.numel() * 4}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if a.dtype == "bf16" else "f32", "data": "synthetic",
            "config": make_config(B, S, a.dtype, world),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "peaks": peak_src,
            "launch_mode": ("cuda graph replay (one graph per drop-bit pattern, %d captured in the warm-up; gpu_launches = kernels per replay)" % len(net._graphs))
                           if use_graph else "eager launches"}
    if world > 1:
        line["rank_ms_per_step"] = [round(v, 3) for v in rank_ms_train]      # ms_per_step is their maximum
        if os.environ.get("SHM_DP_NOREDUCE"):
            line["diagnostic"] = "SHM_DP_NOREDUCE=1: gradient all-reduce skipped (replicas diverge) -- isolates straggler skew from communication"

    if not a.no_extras:
        # ---- inference img/s at N GPUs (BASELINE.json metric; configs[3] shape: SpecSeg mask + generator, batch 64 at 512 x 512):
        # inference is embarrassingly parallel per image -> N replicas, no communication; max-over-ranks device time
        try:
            ib, isz = 64, 512
            inet = M.ShmGANwithSSpecSeg(M.default_args(image_size=isz, batch_size=ib), dtype=a.dtype, allow_random_specseg=True).build()
            inet.cuda_graph = os.environ.get("SHM_CUDA_GRAPH", "1") != "0"
            img = torch.rand((ib, isz, isz, 3), device="cuda")
            for _ in range(3):
                inet.inference_step(img)
            ms_inf = timed(lambda: inet.inference_step(img), 5)
            line["inference_replicas"] = {"images_per_s": world * ib * 5 / (ms_inf * 1e-3), "ms_per_batch": ms_inf / 5, "batch_per_gpu": ib,
                                          "size": isz, "n_gpus": world, "scaling": "replicas only (no collective)"}
            del inet, img
            torch.cuda.empty_cache()
        except Exception as ex:                              # report, do not hide
            line["inference_replicas"] = {"error": str(ex)[:200]}
        trace("inference replicas done")

    if not a.no_extras and world in (2, 4):
        # ---- configs[2] AS WRITTEN: global batch 128 split over the ranks (64 / 32 per GPU; at 8 GPUs the headline line above already is
        # 16 per GPU = global 128).  Strong scaling: total work fixed as N grows.
        try:
            gb = 128
            pb = gb // world
            del net
            torch.cuda.empty_cache()
            net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=pb), dtype=a.dtype, allow_random_specseg=True).build()
            net.enable_data_parallel()
            pol2 = [torch.rand((pb, S, S, 3), generator=g, device="cuda") for _ in range(4)]
            in2 = pol2 + [net.calculate_estimate_diffuse(*pol2)]
            for _ in range(3):
                net.train_step(*in2)
            ms2 = timed(lambda: net.train_step(*in2), 5)
            line["configs2_global_batch_128"] = {"value": gb * 5 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / 5, "batch_per_gpu": pb,
                                                 "global_batch": gb, "n_gpus": world, "scaling": "strong", "steps": 5, "warmup": 3}
            del pol2, in2
        except Exception as ex:
            line["configs2_global_batch_128"] = {"error": str(ex)[:200]}
        del net
        torch.cuda.empty_cache()
        net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype=a.dtype, allow_random_specseg=True).build()
        net.enable_data_parallel()
        for _ in range(2):
            net.train_step(*dev_in)
        trace("strong-scaling leg done")

    if not a.no_extras:
        # ---- two profiled steps for the per-kernel roofline.  EVERY rank steps (the data-parallel gradient all-reduce needs all of
        # them -- a rank-0-only step would wait for its peers forever); only rank 0 records the per-launch CUDA events.
        if rank == 0:
            ops.PROF = []
        net.cuda_graph = False                               # per-launch CUDA events: eager launches on every rank
        net.overlap = False                                  # per-launch events need the kernels of one step serialised on one stream
        for _ in range(2):
            step_resident()
        torch.cuda.synchronize()
        net.overlap = True
        trace("profiled steps done")

    if rank == 0 and not a.no_extras:
        # ---- roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions), CUDA events per launch
        rows = [(fam, kind, name, fl, nb, e0.elapsed_time(e1)) for fam, kind, name, fl, nb, e0, e1 in ops.PROF]
        ops.PROF = None
        step_ms = ms / a.steps
        fam, kern = {}, {}
        for f, kind, name, fl, nb, t in rows:
            kname = f[3:] if f.startswith("tc:") else None
            f0 = "tc" if kname else f
            c = fam.setdefault((f0, kind), [0, 0.0, 0.0, 0])
            c[0] += fl; c[1] += t; c[2] += nb; c[3] += 1
            if kname:
                c = kern.setdefault(kname, [0, 0.0, 0.0, 0])
                c[0] += fl; c[1] += t; c[2] += nb; c[3] += 1
        tc_fl = sum(v[0] for k, v in fam.items() if k[0] == "tc") / 2
        tc_ms = sum(v[1] for k, v in fam.items() if k[0] == "tc") / 2
        # measured DRAM traffic of single launches (ncu --set full, profiles/r0N_ncu_kernels.json), reported beside the live numbers
        # one `ncu --set full` capture per kernel family (profiles/r02_ncu_kernels.json: tools/prof_kernels.py + tools/ncu_summary2.py; the first
        # entry of a family is its representative 256 x 256-step layer)
        ncu = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r02_ncu_kernels.json")) as fh:
                for label, d in json.load(fh).items():
                    ncu.setdefault(d.get("family"), d)
        except Exception:
            ncu = {}
        ktab = {}
        for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1]):
            ktab[k] = {"launches_per_step": v[3] // 2, "ms_per_step": v[1] / 2, "share_of_step": v[1] / 2 / step_ms,
                       "tflops": v[0] / (v[1] * 1e-3) / 1e12, "frac_of_sustained_peak": v[0] / (v[1] * 1e-3) / 1e12 / tf_sus,
                       "avg_launch_ms": v[1] / v[3], "ncu": ncu.get(k)}
        if ktab:
            top = next(iter(ktab))
            t = ktab[top]
            line["roofline"] = {"bound": "tensor", "kernel": top + " (dominant tcgen05 kernel: all its launches in the live step, CUDA events)",
                                # per-launch CUDA events around serialised launches time the kernel ALONE (bursts between bandwidth kernels, at
                                # burst clocks): the burst peak is the ceiling -- against the sustained peak the kernel reads 1.10
                                "achieved": t["tflops"], "peak": tf_burst, "unit": "TFLOP/s", "frac": t["tflops"] / tf_burst,
                                "frac_of_sustained_peak": t["frac_of_sustained_peak"],
                                "traffic": (ncu.get(top) or {}).get("dram_bytes_per_launch"),
                                "traffic_note": (ncu.get(top) or {}).get("note"),
                                "launches_per_step": t["launches_per_step"], "avg_launch_ms": t["avg_launch_ms"],
                                "ms_per_step_in_kernel": t["ms_per_step"], "share_of_step": t["share_of_step"],
                                "peak_kind": "bf16_tflops_burst (%s): each launch timed alone by its own CUDA events" % peak_src}
            line["tc_kernels"] = ktab
            line["tc_family"] = {"achieved": tc_fl / (tc_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "frac": tc_fl / (tc_ms * 1e-3) / 1e12 / tf_sus,
                                 "ms_per_step": tc_ms, "share_of_step": tc_ms / step_ms,
                                 "note": "all tcgen05 conv launches (fwd + dgrad + wgrad), algorithmic FLOPs of the unpadded layers"}
        line["kernel_families"] = {"%s_%s" % k: {"ms_per_step": v[1] / 2, "tflops": (v[0] / (v[1] * 1e-3) / 1e12) if v[1] > 0 else None,
                                                  "gbs_algorithmic": (v[2] / (v[1] * 1e-3) / 1e9) if v[1] > 0 else None,
                                                  "launches": v[3] // 2} for k, v in sorted(fam.items())}
        nb_norm = sum(v[2] for k, v in fam.items() if k[0] == "norm") / 2
        ms_norm = sum(v[1] for k, v in fam.items() if k[0] == "norm") / 2
        if ms_norm > 0:
            line["hbm_family"] = {"achieved": nb_norm / (ms_norm * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nb_norm / (ms_norm * 1e-3) / 1e9 / hbm,
                                  "ms_per_step": ms_norm, "share_of_step": ms_norm / step_ms,
                                  "note": "all instance-norm / activation-backward launches of the live step (cp.async-pipelined range kernels), "
                                          "algorithmic bytes / CUDA-event time"}
        line["conv_ms_per_step"] = sum(r[5] for r in rows if r[0] != "norm") / 2
        line["norm_ms_per_step"] = sum(r[5] for r in rows if r[0] == "norm") / 2
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            per = {}
            for f, kind, name, fl, nb, t in rows:
                c = per.setdefault("%s/%s/%s" % (f, kind, name), [0, 0.0, 0])
                c[0] += fl; c[1] += t; c[2] += 1
            with open(os.path.join(ROOT, "gpurun_out", "bench_conv_layers.json"), "w") as fh:
                json.dump({k: {"launches_per_step": v[2] // 2, "ms_per_step": v[1] / 2, "tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None}
                           for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}, fh, indent=1)
        except OSError:
            pass

        # ---- HBM-bound leg: the pseudo-diffuse min-of-4 kernel at the step's size (4 reads + 1 write)
        n_bytes = 5 * dev_in[0].numel() * 4
        big = [torch.rand((64, S, S, 3), device="cuda") for _ in range(4)]
        for _ in range(3):
            ops.pseudo_diffuse_min4(*big)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.pseudo_diffuse_min4(*big)
        e1.record()
        torch.cuda.synchronize()
        pd_bytes = 5 * big[0].numel() * 4
        gbs = pd_bytes * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        line["roofline_hbm"] = {"bound": "hbm", "kernel": "min4_f32_kernel (pseudo-diffuse, 64x%dx%dx3 fp32 x 4 in + 1 out = %.0f MB > L2)" % (S, S, pd_bytes / 1e6),
                                "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None}
        del big

        # ---- inference img/s (BASELINE.json metric, configs[3] / configs[4] shapes): SpecSeg mask + generator forward + yuv->rgb
        # (test.py:218-250) through inference_step; inputs resident, CUDA events, >= L2-sized tensors.
        # The rank-0-only legs from here on (inference sweep, fp32 parity mode, CPU baseline) run at N = 1 only; at N > 1 the
        # all-rank `inference_replicas` leg above carries the inference number.
        inf = {}
        for tag, (ib, isz) in ({"b64_512": (64, 512), "b8_1024": (8, 1024), "b64_256": (64, 256)} if world == 1 else {}).items():
            try:
                inet = M.ShmGANwithSSpecSeg(M.default_args(image_size=isz, batch_size=ib), dtype=a.dtype, allow_random_specseg=True).build()
                inet.cuda_graph = os.environ.get("SHM_CUDA_GRAPH", "1") != "0"   # replayed batches (the per-launch profile pass below runs eagerly)
                img = torch.rand((ib, isz, isz, 3), device="cuda")
                for _ in range(3):
                    inet.inference_step(img)
                torch.cuda.synchronize()
                # three timed groups of 5 batches, best group: the first group after a cache flush still pays for the allocator growing its pools
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5):
                        inet.inference_step(img)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) / 5)
                t = min(ts)
                gf = 119.5 * (isz / 256.0) ** 2              # SpecSeg + G1 with the live mask branch, GFLOP per image
                inf[tag] = {"images_per_s": ib / (t * 1e-3), "ms_per_batch": t, "batch": ib, "size": isz,
                            "tflops": gf * ib / t}
                # where the batch's time goes: per-launch CUDA events of one more pass, by kernel family
                ops.PROF = []
                inet.inference_step(img)
                torch.cuda.synchronize()
                fams = {}
                for f, kind, name, fl, nb, ev0, ev1 in ops.PROF:
                    c = fams.setdefault(("tc" if f.startswith("tc:") else f) + "_" + kind, [0.0, 0.0, 0.0, 0])
                    c[0] += ev0.elapsed_time(ev1); c[1] += fl; c[2] += nb; c[3] += 1
                ops.PROF = None
                inf[tag]["families"] = {k: {"ms": v[0], "tflops": v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else None,
                                            "gbs": v[2] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None, "launches": v[3]}
                                        for k, v in sorted(fams.items(), key=lambda kv: -kv[1][0])}
                del inet, img
                torch.cuda.empty_cache()
            except Exception as ex:                          # report, do not hide
                inf[tag] = {"error": str(ex)[:200]}
        if world == 1:
            line["inference"] = inf
            if "images_per_s" in inf.get("b64_512", {}):
                line["inference_512_images_per_s"] = inf["b64_512"]["images_per_s"]      # configs[3] headline, short enough to survive any tail
            # ---- configs[0]: the reference's CPU-runnable case beside the inference legs -- SpecSeg mask + random-init generator forward
            # on 1 x 256 x 256 x 3 fp32 (test.py:218-250) through the oracle port on every host core (TensorFlow is not installable)
            try:
                import oracle as O
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                f32 = torch.float32
                Gp = O.init_params(O.generator_param_specs(64, True), 42, f32)
                Sp = O.init_params(O.specseg_param_specs(), 44, f32)
                img1 = torch.rand((1, 256, 256, 3), generator=torch.Generator().manual_seed(0))
                with torch.no_grad():
                    O.inference_step(Gp, Sp, img1)
                    t0 = time.perf_counter()
                    reps = 5
                    for _ in range(reps):
                        O.inference_step(Gp, Sp, img1)
                    dt1 = (time.perf_counter() - t0) / reps
                line["inference_cpu_baseline"] = {"value": 1.0 / dt1, "unit": "images/s", "cores": cores, "kind": "port", "ms_per_image": dt1 * 1e3,
                                                  "sample": "configs[0]: SpecSeg mask + generator forward + yuv->rgb on 1x256x256x3 fp32, PyTorch-CPU oracle "
                                                            "port, %d host threads, mean of %d images" % (cores, reps),
                                                  "gpu_b64_256_images_per_s": inf.get("b64_256", {}).get("images_per_s")}
            except Exception as ex:
                line["inference_cpu_baseline"] = {"error": str(ex)[:200]}

        # ---- configs[1] literally: the fp32 parity mode on the same batch (2 steps)
        if a.dtype == "bf16" and world == 1:
            del net
            torch.cuda.empty_cache()
            net32 = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="fp32", allow_random_specseg=True).build()
            net32.train_step(*dev_in)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                net32.train_step(*dev_in)
            torch.cuda.synchronize()
            line["parity_mode_fp32"] = {"value": 2 * B / (time.perf_counter() - t0), "unit": UNIT, "steps": 2,
                                        "note": "configs[1]: exact-fp32 SIMT kernels, same batch"}
            del net32
            torch.cuda.empty_cache()

        # ---- CPU baseline beside it (rank 0, N = 1 only): one sample of the same step through the oracle port on the host cores
        if world == 1:
            fn, cores = cpu_train_step_fn(S)
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "1 train step of ONE sample (B=1) at %dx%d, float32, PyTorch-CPU oracle port of the reference step "
                                              "(TensorFlow not installable), %d host threads, %.1f s" % (S, S, cores, dt)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    trace("printing / leaving")
    if world > 1:
        dist.barrier()                                       # leave together: rank 0's bookkeeping above has no collective in it
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
