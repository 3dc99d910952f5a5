"""On-GPU data-parallel equivalence: two ranks of `model.train_step`, each on its shard of a global batch, must
end every step with BIT-IDENTICAL parameters, and those must equal a single-rank step on the concatenated batch up to the fp32 summation
order (weight-gradient atomics, the all-reduce) -- in the fp32 parity mode and in the bf16 tensor-core mode, over two optimiser steps
(so that the overlapped, range-by-range all-reduce of the generator's gradients and the version bump after the broadcast both matter).

Two processes are spawned.  With >= 2 visible GPUs each rank takes its own device and the collective is NCCL; on a single-GPU box both
ranks share cuda:0 and the process group is gloo over CUDA tensors (NCCL refuses two ranks on one device) -- GradReducer is backend-agnostic,
so the step being tested is the same."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

S = 64
BITS, T = [True, False, True, False, False], 0.9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(GB, fs):
    g = torch.Generator().manual_seed(77)
    pol = [torch.rand((GB, S, S, 3), generator=g) for _ in range(4)]
    ed = torch.minimum(torch.minimum(pol[0], pol[1]), torch.minimum(pol[2], pol[3]))
    noise = (torch.randn((2, GB, S, S, 3), generator=g) * 0.1).bfloat16().float()       # [training=True call][sample]
    keep = (torch.rand((2, GB, S // 32, S // 32, fs * 16), generator=g) < 0.8).float()
    return pol + [ed], noise, keep


def _fresh(dtype, fs, B):
    from shmgan_b200 import model as M
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype=dtype, allow_random_specseg=True).build()
    net.drop_bits, net.TARGET_LABELS = BITS, T
    return net


def _steps(net, origs, noise, keep, lo, hi, nsteps):
    """Runs nsteps train steps on samples [lo, hi); returns (losses per step, gradient / parameter snapshots after the FIRST step)."""
    dev = lambda t: t.cuda().contiguous()
    net.d_noise = dev(torch.cat([noise[0, lo:hi], noise[1, lo:hi]]))
    net.d_keep = dev(torch.cat([keep[0, lo:hi], keep[1, lo:hi]]))
    batch = [dev(o[lo:hi]) for o in origs]
    losses, snap = [], None
    for i in range(nsteps):
        net.train_step(*batch)
        losses.append((net.total_Generator_loss, net.total_Discriminator_loss))
        if i == 0:
            snap = [t.clone() for t in (net.G.net.store.grad, net.D.net.store.grad, net.G.net.store.flat, net.D.net.store.flat)]
    torch.cuda.synchronize()
    return losses, snap


def _worker(rank, world, port, dtype, fs, GB, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    multi = torch.cuda.device_count() >= world
    torch.cuda.set_device(rank if multi else 0)
    dist.init_process_group("nccl" if multi else "gloo", rank=rank, world_size=world)
    try:
        origs, noise, keep = _inputs(GB, fs)
        per = GB // world
        net = _fresh(dtype, fs, per)
        if rank == 1:                                                 # rank 1 starts from DIFFERENT weights and has already run a forward
            for st in (net.G.net.store, net.D.net.store, net.SpecSeg.net.store):
                st.init(900 + rank)
            _steps(net, origs, noise, keep, 0, per, 1)                # ... so its bf16 weight copies are stale when the broadcast lands
            for st in (net.G.net.store, net.D.net.store):
                st.m.zero_(); st.v.zero_(); st.step = 0
            net.step_count = 0
        net.enable_data_parallel()
        losses, snap = _steps(net, origs, noise, keep, rank * per, (rank + 1) * per, 2)
        flats = [net.G.net.store.flat, net.D.net.store.flat]
        same = True
        for f in flats:
            both = [torch.empty_like(f) for _ in range(world)]
            dist.all_gather(both, f)
            same = same and all(torch.equal(both[0], b) for b in both[1:])
        res = {"rank": rank, "identical": bool(same), "backend": "nccl" if multi else "gloo", "bytes": net._reducer.bytes_reduced}
        if rank == 0:
            del net
            torch.cuda.empty_cache()
            ref = _fresh(dtype, fs, GB)                                # one rank, the whole global batch, no reducer
            ref_losses, rsnap = _steps(ref, origs, noise, keep, 0, GB, 2)
            rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
            # step 1: the gradient buffers hold the SUM over ranks (clip + Adam apply 1 / world)
            res["grad_err"] = [rel(snap[0] / world, rsnap[0]), rel(snap[1] / world, rsnap[1])]
            cos = lambda a, b: float((a.double() @ b.double()) / (a.double().norm() * b.double().norm()))
            res["grad_cos"] = [cos(snap[0], rsnap[0]), cos(snap[1], rsnap[1])]
            dG, dD = (snap[2] - rsnap[2]).abs(), (snap[3] - rsnap[3]).abs()
            res["param_max_diff"] = [float(dG.max()), float(dD.max())]
            res["param_frac_moved"] = [float((dG > 1e-6).float().mean()), float((dD > 1e-6).float().mean())]
            res["ref_loss"] = [list(l) for l in ref_losses]
        loss_t = torch.tensor([v for l in losses for v in l], dtype=torch.float64, device="cuda")
        dist.all_reduce(loss_t)
        res["mean_loss"] = (loss_t / world).tolist()
        out.put(res)
    finally:
        dist.destroy_process_group()


# fp32 parity mode: shard and global runs differ only in fp32 summation order.  bf16 mode: the per-sample forward is NOT bit-identical between a
# batch of 8 and a batch of 16 (instance-norm partial sums are grouped by CTA, so a statistic moves in its last bit, a bf16 rounding flips here
# and there, and the random-init networks amplify that exactly as they amplify bf16 storage noise -- tests/test_gpu_nets.py measures 0.1-0.35
# per tensor for that); measured here: 0.7 % (G) / 6 % (D) max-norm on the whole flat buffer.  Bound: 0.15 max-norm and cosine >= 0.99.
@pytest.mark.parametrize("dtype,fs,GB,gtol,gcos", [("fp32", 8, 4, 2e-4, 0.999999), ("bf16", 64, 16, 0.15, 0.99)])
def test_two_ranks_equal_one_rank_on_the_global_batch(dtype, fs, GB, gtol, gcos):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dtype, fs, GB, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0
    res = {r["rank"]: r for r in (q.get() for _ in range(2))}
    r0 = res[0]
    assert r0["identical"] and res[1]["identical"], "ranks diverged"
    assert r0["bytes"] > 0
    # step 1: averaged all-reduced gradients == gradients of the global batch (same per-sample arithmetic; only the fp32 summation order of
    # the weight-gradient atomics and of the all-reduce differs)
    assert max(r0["grad_err"]) < gtol and min(r0["grad_cos"]) > gcos, r0
    # parameters after the first clip + Adam step: Adam's first step moves every weight by ~lr_t * g / |g| (~2e-5), so a gradient whose sign
    # is inside summation noise may move the other way: allowed for a vanishing fraction, bounded by two full steps
    assert max(r0["param_max_diff"]) < 1e-4 and max(r0["param_frac_moved"]) < (1e-3 if dtype == "fp32" else 0.2), r0
    means, want = r0["mean_loss"], [v for l in r0["ref_loss"] for v in l]
    ltol = 1e-4 if dtype == "fp32" else 5e-3
    assert means[0] == pytest.approx(want[0], rel=ltol) and means[1] == pytest.approx(want[1], rel=ltol), r0      # step 1: shard means average
    try:
        import json
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "dp_equivalence_%s.json" % dtype), "w") as fh:
            json.dump(res, fh, indent=1)
    except OSError:
        pass
    # step 2 starts from (almost) the same weights; the random-init networks amplify the few sign-level differences, so only track the loss
    assert means[2] == pytest.approx(want[2], rel=5e-2) and means[3] == pytest.approx(want[3], rel=5e-2), r0
