"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/shmgan.h declares, argument validation
fails loudly without touching a device, and the data-parallel host logic (sharding, bucketing, the gradient reducer) is
exercised with world_size = 2 over the gloo backend."""
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = ""
    for name in sorted(os.listdir(os.path.join(ROOT, "include"))):       # shmgan.h (the boundary) and shmgan_tools.h (measurement hooks)
        if name.endswith(".h"):
            with open(os.path.join(ROOT, "include", name)) as f:
                text += re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(shm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from shmgan_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else _lib.load()
    names = _declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes table the host code binds with covers the same set
    assert not [n for n in names if n not in _lib.EXPORTS], [n for n in names if n not in _lib.EXPORTS]


def test_bad_arguments_fail_loudly_without_a_device():
    import ctypes as C
    from shmgan_b200 import _lib as L
    lib = L.load()
    assert lib.shm_version() >= 1
    d = L.ConvDesc(1, 8, 8, 4, 4, 5, 5, 1, 0, 0, 4, 4, 0, 0)          # 5x5 kernel
    with pytest.raises(L.ShmError):
        L.call("shm_conv2d_fwd", C.byref(d), None, None, None, None, None)
    with pytest.raises(L.ShmError):
        L.call("shm_pseudo_diffuse_min4", None, None, None, None, None, 16, 0, None)
    assert isinstance(lib.shm_last_error(), bytes) and len(lib.shm_last_error()) > 0


def test_shard_bounds_and_buckets():
    from shmgan_b200.parallel import bucket_ranges, shard_bounds
    assert [shard_bounds(128, r, 8) for r in (0, 3, 7)] == [(0, 16), (48, 64), (112, 128)]
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
    r = bucket_ranges(1000, 256)
    assert r[0] == (0, 256) and r[-1][1] == 1000 and all(a % 4 == 0 for a, _ in r)
    assert sum(b - a for a, b in r) == 1000
    assert bucket_ranges(10, 1 << 20, lo=4) == [(4, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reducer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shmgan_b200.parallel import GradReducer
        red = GradReducer(None, bucket_mb=0.001)                       # 262-element buckets -> several collectives
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(1003, generator=g)
        mine = flat.clone()
        params = torch.full((17,), float(rank))
        red.broadcast_params([params])
        red.reduce_async(flat)
        red.wait()
        others = [torch.randn(1003, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        want = sum(others)
        ok = torch.allclose(flat, want, atol=1e-6) and float(params.abs().max()) == 0.0 and red.bytes_reduced == 1003 * 4
        ok = ok and not torch.equal(mine, flat)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_grad_reducer_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get() for _ in range(2))
    assert got == {0: True, 1: True}


def _dp_worker(rank, world, port, out):
    """Each rank runs the oracle's train step on its shard of a global batch of 2; the averaged (all-reduced / world)
    gradients must equal the oracle's gradients on the whole batch: the property the data-parallel step relies on
    (per-image statistics, batch-mean losses)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shmgan_b200.parallel import GradReducer, shard_bounds
        f, S, fs, GB = torch.float64, 32, 4, 2
        Gp = O.init_params(O.generator_param_specs(fs, True), 1, f, randomize_all=True)
        Dp = O.init_params(O.discriminator_param_specs(S, fs, True), 2, f, randomize_all=True)
        g = torch.Generator().manual_seed(0)
        pol = [torch.rand((GB, S, S, 3), generator=g, dtype=f) for _ in range(4)]
        origs = pol + [O.pseudo_diffuse_min4(*pol)]
        mask = torch.rand((GB, S, S, 1), generator=g, dtype=f)
        bits = [True, False, False, True, False]
        a, b = shard_bounds(GB, rank, world)
        _, gG, gD = O.train_step_grads(Gp, Dp, [o[a:b] for o in origs], mask[a:b], bits, 0.9, None, None, True, True, clip=False)
        flat = torch.cat([v.reshape(-1) for v in list(gG.values()) + list(gD.values())]).contiguous()
        red = GradReducer(None, bucket_mb=0.01)
        red.reduce_async(flat)
        red.wait()
        flat /= world
        ok = True
        if rank == 0:
            _, wG, wD = O.train_step_grads(Gp, Dp, origs, mask, bits, 0.9, None, None, True, True, clip=False)
            want = torch.cat([v.reshape(-1) for v in list(wG.values()) + list(wD.values())])
            err = float((flat - want).abs().max() / want.abs().max())
            ok = err < 1e-9
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_average_equals_global_batch_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    got = dict(q.get() for _ in range(2))
    assert got == {0: True, 1: True}


def _range_worker(rank, world, port, out):
    """The generator's gradient buffer is reduced range by range while the backward runs (decoder levels, bottleneck, then the head):
    the ranges must tile the buffer and the result must equal one whole-buffer reduction."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shmgan_b200.parallel import GradReducer
        g = torch.Generator().manual_seed(rank)
        flat = torch.randn(10_000, generator=g, dtype=torch.float64)
        mine = flat.clone()
        red = GradReducer(None, bucket_mb=0.004)                       # 1048-element buckets: several per range
        for lo, hi in ((7000, 10_000), (5200, 7000), (4000, 5200), (0, 4000)):
            red.reduce_async(flat, lo, hi)
        red.wait()
        both = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        out.put((rank, bool(torch.equal(flat, both[0] + both[1])) and red.bytes_reduced == 8 * 10_000))
    finally:
        dist.destroy_process_group()


def test_ranged_bucket_reduce_covers_buffer_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_range_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(q.get() for _ in range(2)) == {0: True, 1: True}


def test_generator_grad_ranges_tile_the_flat_buffer():
    """Generator.grad_range (the ranges the overlapped all-reduce uses) from the parameter inventory alone: decoder levels 3..0, bottleneck,
    head -- contiguous, disjoint, covering exactly the trainable prefix, each range holding exactly the layers whose gradients are final."""
    from shmgan_b200 import nets

    class _Store:                                                      # ParamStore's offset rule without device buffers
        def __init__(self, specs):
            order = [s for s in specs if nets._is_trainable(s[0])] + [s for s in specs if not nets._is_trainable(s[0])]
            self.offsets, off = {}, 0
            for name, shape, _ in order:
                n = math.prod(shape)
                self.offsets[name] = (off, n, shape)
                off += (n + 3) // 4 * 4
                if nets._is_trainable(name):
                    self.n_train = off

    import math
    fake = type("G", (), {"store": _Store(nets.generator_specs(64, True)), "grad_range": nets.Generator.grad_range})()
    ranges = [fake.grad_range(("dec", u)) for u in (3, 2, 1, 0)] + [fake.grad_range("bott"), fake.grad_range("head")]
    assert ranges[0][1] == fake.store.n_train and ranges[-1][0] == 0
    for (lo, hi), (lo2, hi2) in zip(ranges[1:], ranges[:-1]):
        assert hi == lo2 and lo < hi                                   # each range ends where the previous (later-layer) one starts
    off = fake.store.offsets
    lo, hi = fake.grad_range(("dec", 3))
    assert lo == off["up4T.w"][0] and off["out.b"][0] + 4 == hi        # the output layer rides with the last decoder level
    lo, hi = fake.grad_range("head")
    assert all(lo <= off[k][0] < hi for k in ("enc1a.w", "attn4b.b", "enc4b.b")) and off["bott1.w"][0] == hi
