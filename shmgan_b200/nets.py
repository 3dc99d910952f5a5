"""The three networks of the SHMGAN hot path, orchestrated layer by layer over libshmgan kernels.

  Generator      build_generator       ShmGANwithSSpecSeg.py:228-327  (+ attention_layer :404-412, live mask)
  Discriminator  build_discriminator   ShmGANwithSSpecSeg.py:343-389
  SpecSeg        SpecSeg               SpecSeg.py:27-98 (predict only)

Parameters live in ONE flat fp32 buffer per network (Keras creation order, Keras layouts) with matching flat
gradient / Adam m / Adam v buffers, so clip+Adam is a single launch and the data-parallel all-reduce works on
contiguous buckets.  Forward passes save exactly what the hand-written backward needs (post-activation conv outputs,
instance-norm sums, layer inputs); concat is never materialised: producers write into channel slices.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, Conv


# ------------------------------------------------------------------------------------------------
# parameter inventory (names / shapes / initialisers follow the reference layer by layer)
# ------------------------------------------------------------------------------------------------
def generator_specs(filter_size=64, live_mask=True):
    """[(name, shape, kind)].  kind: w N(0,.02) (:200) | b zeros | g IN gamma = 1 | be IN beta N(0,.02) (frozen, Q2)."""
    specs, N, cin = [], filter_size, 10
    for lvl in range(1, 5):
        for ab in "ab":
            specs += [(f"enc{lvl}{ab}.w", (3, 3, cin, N), "w"), (f"enc{lvl}{ab}.b", (N,), "b"),
                      (f"enc{lvl}{ab}.in_gamma", (N,), "g"), (f"enc{lvl}{ab}.in_beta", (N,), "be")]
            cin = N
        if live_mask:
            specs += [(f"attn{lvl}a.w", (3, 3, 1, N), "w"), (f"attn{lvl}a.b", (N,), "b"),
                      (f"attn{lvl}b.w", (3, 3, N, N), "w"), (f"attn{lvl}b.b", (N,), "b")]
        if lvl < 4:
            N *= 2
    for i in (1, 2):
        specs += [(f"bott{i}.w", (1, 1, N, N), "w"), (f"bott{i}.b", (N,), "b"),
                  (f"bott{i}.in_gamma", (N,), "g"), (f"bott{i}.in_beta", (N,), "be")]
    cin = N
    for u in range(1, 5):
        if u > 1:
            N //= 2
        specs += [(f"up{u}T.w", (3, 3, N, cin), "w"), (f"up{u}T.b", (N,), "b")]
        cin = 2 * N
        for ab in "ab":
            specs += [(f"dec{u}{ab}.w", (3, 3, cin, N), "w"), (f"dec{u}{ab}.b", (N,), "b"),
                      (f"dec{u}{ab}.in_gamma", (N,), "g"), (f"dec{u}{ab}.in_beta", (N,), "be")]
            cin = N
    specs += [("out.w", (1, 1, cin, 1), "w"), ("out.b", (1,), "b")]
    return specs


def discriminator_specs(image_size, filter_size=64, live_mask=True):
    specs, N, cin = [], filter_size, 3
    for i, mult in enumerate((1, 2, 4, 8), start=1):
        specs += [(f"d{i}.w", (3, 3, cin, N * mult), "w"),
                  (f"d{i}.in_gamma", (N * mult,), "g"), (f"d{i}.in_beta", (N * mult,), "be")]
        cin = N * mult
    if live_mask:
        specs += [("dattn_a.w", (3, 3, 1, cin), "w"), ("dattn_a.b", (cin,), "b"),
                  ("dattn_b.w", (3, 3, cin, cin), "w"), ("dattn_b.b", (cin,), "b")]
    specs += [("d5.w", (3, 3, cin, N * 16), "w"), ("d5.in_gamma", (N * 16,), "g"), ("d5.in_beta", (N * 16,), "be")]
    cin = N * 16
    specs += [("head.w", (3, 3, cin, 1), "w")]
    s32 = image_size // 32
    specs += [("dense.w", (s32 * s32 * cin, 5), "w")]
    return specs


def specseg_specs():
    specs, cin = [], 1
    for i, c in enumerate((16, 32, 64, 128, 256), start=1):
        specs += [(f"c{i}a.w", (3, 3, cin, c), "w5"), (f"c{i}a.b", (c,), "b"),
                  (f"c{i}b.w", (3, 3, c, c), "w5"), (f"c{i}b.b", (c,), "b"),
                  (f"bn{i}.gamma", (c,), "g"), (f"bn{i}.beta", (c,), "b"),
                  (f"bn{i}.mean", (c,), "b"), (f"bn{i}.var", (c,), "g")]
        cin = c
    for i, c in zip((6, 7, 8, 9), (128, 64, 32, 16)):
        specs += [(f"u{i}.w", (2, 2, c, cin), "glorot"), (f"u{i}.b", (c,), "b"),
                  (f"c{i}a.w", (3, 3, 2 * c, c), "w5"), (f"c{i}a.b", (c,), "b"),
                  (f"c{i}b.w", (3, 3, c, c), "w5"), (f"c{i}b.b", (c,), "b")]
        cin = c
    specs += [("out.w", (1, 1, 16, 1), "glorot"), ("out.b", (1,), "b")]
    return specs


def _is_trainable(name: str) -> bool:
    return not (name.endswith("in_gamma") or name.endswith("in_beta") or name.startswith("bn"))


class ParamStore:
    """Flat fp32 parameter buffer with named views; trainable entries first (so grads / Adam state cover a prefix)."""

    def __init__(self, specs, device="cuda", seed=42, trainable=True):
        order = [s for s in specs if _is_trainable(s[0])] + [s for s in specs if not _is_trainable(s[0])]
        self.specs = specs
        self.offsets: Dict[str, tuple] = OrderedDict()
        off = 0
        for name, shape, _ in order:
            n = math.prod(shape)
            self.offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4                       # keep every view 16-byte aligned
            if _is_trainable(name):
                self.n_train = off
        self.total = off
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = OrderedDict(
            (k, self.flat[o:o + n].view(shape)) for k, (o, n, shape) in self.offsets.items())
        self.grad = self.m = self.v = None
        self.gviews: Dict[str, torch.Tensor] = {}
        if trainable:
            self.grad = torch.zeros(self.n_train, dtype=torch.float32, device=device)
            self.m = torch.zeros_like(self.grad)
            self.v = torch.zeros_like(self.grad)
            self.gviews = OrderedDict((k, self.grad[o:o + n].view(shape)) for k, (o, n, shape) in self.offsets.items()
                                      if _is_trainable(k))
        self.version = 0
        self.tc_convs: List[Conv] = []
        self._tc_jobs = None
        self.step = 0
        self.padded_convs: List[Conv] = []
        self.init(seed)

    def init(self, seed):
        """Seeded reference initialisers (host RNG; one H2D copy)."""
        g = torch.Generator().manual_seed(seed)
        host = torch.zeros(self.total, dtype=torch.float32)
        for name, shape, kind in self.specs:
            o, n, _ = self.offsets[name]
            if kind == "w":
                t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.02
            elif kind == "w5":
                t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.05
            elif kind == "glorot":
                rf = math.prod(shape[:-2])
                fan_in, fan_out = shape[-1] * rf, shape[-2] * rf
                if shape[0] == 1:
                    fan_in, fan_out = shape[-2], shape[-1]
                t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * math.sqrt(6.0 / (fan_in + fan_out))
            elif kind == "be":
                t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.02
            elif kind == "g":
                t = torch.ones(shape, dtype=torch.float64)
            else:
                t = torch.zeros(shape, dtype=torch.float64)
            host[o:o + n] = t.reshape(-1).float()
        self.flat.copy_(host)
        self.version += 1

    def load(self, named: Dict[str, torch.Tensor]):
        """Imports parameters given in the reference (Keras) layouts, e.g. the oracle's OrderedDict."""
        host = self.flat.cpu()
        for k, t in named.items():
            o, n, shape = self.offsets[k]
            assert tuple(t.shape) == tuple(shape), (k, t.shape, shape)
            host[o:o + n] = t.detach().reshape(-1).float().cpu()
        self.flat.copy_(host)
        self.version += 1

    def export(self) -> Dict[str, torch.Tensor]:
        host = self.flat.cpu()
        return OrderedDict((k, host[o:o + n].view(shape).clone()) for k, (o, n, shape) in self.offsets.items())

    def export_grads(self) -> Dict[str, torch.Tensor]:
        self.finalize_grads()
        host = self.grad.cpu()
        return OrderedDict((k, host[o:o + n].view(shape).clone()) for k, (o, n, shape) in self.offsets.items() if _is_trainable(k))

    def bind(self, conv: Conv):
        conv.store = self
        conv.w = self.views[conv.name + ".w"]
        conv.dw = self.gviews.get(conv.name + ".w")
        if conv.has_bias:
            conv.b = self.views[conv.name + ".b"]
            conv.db = self.gviews.get(conv.name + ".b")
        if conv.cin_pad is not None and conv not in self.padded_convs:
            self.padded_convs.append(conv)
        return conv

    def pad(self, conv: Conv):
        conv.enable_pad()
        self.padded_convs.append(conv)
        return conv

    # -- batched refresh of the bf16 tensor-core weight copies --------------------------------------------------------
    def register_tc(self, conv: Conv):
        if conv not in self.tc_convs and type(conv) is Conv:
            self.tc_convs.append(conv)
            self._tc_jobs = None

    def refresh_tc_all(self):
        """ONE launch re-lays every registered layer's weights (fwd + dgrad layouts) after an optimiser step; layers seen for the
        first time are refreshed individually by Conv.refresh_tc and join the batch from the next step on."""
        todo = [c for c in self.tc_convs if c.tc_version != self.version]
        if len(todo) < 2:
            return
        if self._tc_jobs is None or self._tc_jobs[0] != len(self.tc_convs):
            blob = bytearray(b"".join(c.prep_job() for c in self.tc_convs))
            import ctypes as C
            host = (C.c_char * len(blob)).from_buffer(blob)
            total = int(ops.call("shm_conv2d_tc_prep_jobs_finalize", host, len(self.tc_convs)))
            dev = torch.frombuffer(blob, dtype=torch.uint8).clone().to(self.flat.device)
            self._tc_jobs = (len(self.tc_convs), dev, total)
        n, dev, total = self._tc_jobs
        ops.call("shm_conv2d_tc_prep_multi", ops._p(dev), n, total, ops._stream())
        for c in self.tc_convs:
            c.tc_version = self.version

    def zero_grad(self):
        ops.zero_(self.grad)
        for c in self.padded_convs:
            c.zero_pad_grad()
        self._finalized = False

    def finalize_grads(self):
        """Copies the weight gradients of the zero-padded first layers from their 64-channel scratch into the flat buffer.  Runs ONCE per
        zero_grad(): call it when the backward sweeps are done, before the all-reduce / clip / Adam.  (A second fold after the data-parallel
        all-reduce would overwrite the reduced gradients of those layers with the rank-local ones: the ranks' first layers drifted apart in
        round 1 -- caught by tests/test_gpu_data_parallel.py.)"""
        if getattr(self, "_finalized", False):
            return
        for c in self.padded_convs:
            c.fold_pad_grad()
        self._finalized = True

    def lr_t(self, lr0=2e-5, beta1=0.5, beta2=0.99, decay_steps=10000.0, decay_rate=0.95) -> float:
        """Step size of the NEXT adam_step: ExponentialDecay learning rate x Adam bias correction (ShmGANwithSSpecSeg.py:169-175)."""
        t = self.step + 1
        lr = lr0 * decay_rate ** (self.step / decay_steps)
        return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)

    def adam_step(self, lr0=2e-5, beta1=0.5, beta2=0.99, eps=1e-7, clip=1.0, gscale=1.0,
                  decay_steps=10000.0, decay_rate=0.95, lr_dev=None):
        """clip_by_value(+-1) + Keras Adam with ExponentialDecay (ShmGANwithSSpecSeg.py:169-175, :860-871).
        lr_dev: one-element fp32 CUDA tensor holding lr_t(), read when the kernel runs (a captured step is replayed with the current value)."""
        self.finalize_grads()
        if lr_dev is not None:
            ops.call("shm_clip_adam_dev", ops._p(self.flat), ops._p(self.grad), ops._p(self.m), ops._p(self.v), self.n_train,
                     ops._p(lr_dev), beta1, beta2, eps, clip, gscale, ops._stream())
        else:
            ops.call("shm_clip_adam", ops._p(self.flat), ops._p(self.grad), ops._p(self.m), ops._p(self.v), self.n_train,
                     self.lr_t(lr0, beta1, beta2, decay_steps, decay_rate), beta1, beta2, eps, clip, gscale, ops._stream())
        self.step += 1
        self.version += 1


# ------------------------------------------------------------------------------------------------
# Generator
# ------------------------------------------------------------------------------------------------
class _CLI:
    """Conv -> (+bias) -> LeakyReLU -> InstanceNorm (ShmGANwithSSpecSeg.py:244-245, :386-389)."""

    def __init__(self, store: ParamStore, name, cin, cout, k=3, stride=1, bias=True):
        self.conv = store.bind(Conv(name, k, k, cin, cout, stride=stride, act=ACT_LRELU, bias=bias))
        self.gamma = store.views[name + ".in_gamma"]
        self.beta = store.views[name + ".in_beta"]


class Generator:
    def __init__(self, filter_size=64, live_mask=True, dtype=torch.float32, device="cuda", seed=42, tensor_core=True):
        self.N0, self.live_mask, self.dtype, self.tc = filter_size, live_mask, dtype, tensor_core
        # tensor-core mode: the 10-channel input and the 1-channel mask are presented zero-padded to 64 channels
        self.pad_in = bool(tensor_core and dtype == torch.bfloat16 and filter_size % 64 == 0)
        self.cin_buf = 64 if self.pad_in else 10
        self.store = ParamStore(generator_specs(filter_size, live_mask), device, seed)
        s, N, cin = self.store, filter_size, 10
        self.enc, self.attn, self.dec, self.up = [], [], [], []
        for lvl in range(1, 5):
            a = _CLI(s, f"enc{lvl}a", cin, N)
            b = _CLI(s, f"enc{lvl}b", N, N)
            self.enc.append((a, b))
            if live_mask:
                self.attn.append((s.bind(Conv(f"attn{lvl}a", 3, 3, 1, N)), s.bind(Conv(f"attn{lvl}b", 3, 3, N, N))))
            cin = N
            if lvl < 4:
                N *= 2
        self.bott = [_CLI(s, "bott1", N, N, k=1), _CLI(s, "bott2", N, N, k=1)]
        cin = N
        for u in range(1, 5):
            if u > 1:
                N //= 2
            self.up.append(s.bind(Conv(f"up{u}T", 3, 3, cin, N, stride=2, transposed=True)))
            self.dec.append((_CLI(s, f"dec{u}a", 2 * N, N), _CLI(s, f"dec{u}b", N, N)))
            cin = N
        self.out = s.bind(Conv("out", 1, 1, cin, 1))
        # forward-only (inference) twins of the few-channel first layers on the THIN halo kernel: the 10-channel input / 1-channel mask
        # padded to 16 channels (32-byte pixel rows) instead of 64 -- a quarter of the bytes and of the MMAs.  Training keeps the
        # 64-channel form because its weight-gradient kernels read the same padded tensor.
        self.thin = {}
        if self.pad_in:
            s.pad(self.enc[0][0].conv)
            for ca, _ in self.attn:
                s.pad(ca)
            if filter_size == 64:
                P = ops.PaddedConv
                self.thin["enc1a"] = s.bind(P("enc1a", 3, 3, 10, 64, 10, 1, act=ACT_LRELU, seg_pad=16, cout_dev=64))
                if live_mask:
                    self.thin["attn1a"] = s.bind(P("attn1a", 3, 3, 1, 64, 1, 1, act=ACT_LRELU, seg_pad=16, cout_dev=64))
                    self.thin["attn2a"] = s.bind(P("attn2a", 3, 3, 1, 128, 1, 1, act=ACT_LRELU, seg_pad=16, cout_dev=128))

    # -- helpers -----------------------------------------------------------------------------------
    def in_channels(self, n, h, w, infer: bool = False) -> int:
        """Channel count of the input buffer forward() wants for an [n,h,w] batch: 16 (thin first layer: forward and dgrad on the thin
        halo kernel, weight gradient through a zero-filling TMA box), 64 (zero-padded first layer) when only that form can serve the
        shape, else the plain 10."""
        c = self.enc[0][0].conv
        if "enc1a" in self.thin and self.thin["enc1a"].servable(n, h, w) and (infer or (c.can_pad(n, h, w) and h % 16 == 0 and w % 8 == 0)):
            return 16
        return 64 if (self.pad_in and c.can_pad(n, h, w)) else 10

    def _conv(self, c: Conv, x, y=None, stats=False):
        return c.fwd(x, y, self.tc, self.store.version, want_stats=stats)

    def attention(self, mask: torch.Tensor, infer: bool = False):
        """attention_layer on a live mask: level 1 un-pooled, then MaxPool2 chain; 2 x [Conv3x3 + LeakyReLU] per level.
        Returns (features, saved) ; computed ONCE per step and shared by every generator pass.  infer: no backward will follow
        (the first conv of levels 1-2 then runs on the thin kernel)."""
        feats, saved, pooled = [], [], ops.cast(mask, self.dtype) if mask.dtype != self.dtype else mask
        for lvl in range(4):
            if lvl > 0:
                pooled = ops.maxpool(pooled, 2)
            ca, cb = self.attn[lvl]
            thin = self.thin.get("attn%da" % (lvl + 1))
            if thin is not None and thin.servable(*pooled.shape[:3]) and (infer or ca.can_pad(*pooled.shape[:3])):
                # 16-channel mask input: thin forward; the weight gradient (training) reads the same tensor through a zero-filling TMA box
                pin = ops.pad_channels(pooled, 16)
                a1 = thin.fwd(pin, None, True, self.store.version)
                a2 = self._conv(cb, a1)
                feats.append(a2)
                saved.append((pin, a1, a2))
                continue
            pin = ops.pad64(pooled) if (self.pad_in and ca.can_pad(*pooled.shape[:3])) else pooled
            a1 = self._conv(ca, pin)
            a2 = self._conv(cb, a1)
            feats.append(a2)
            saved.append((pin, a1, a2))
        return feats, saved

    def attention_backward(self, saved, dattn: List[torch.Tensor]):
        """dattn[lvl] = sum over passes of d(skip)."""
        for lvl in range(4):
            pooled, a1, a2 = saved[lvl]
            ca, cb = self.attn[lvl]
            d2 = ops.act_bwd(dattn[lvl], a2, ACT_LRELU, dbias=cb.db)
            cb.wgrad(a1, d2, self.tc, bias_done=True)
            d1 = cb.dgrad(d2, a1.shape, None, self.tc, self.store.version)
            d1 = ops.act_bwd(d1, a1, ACT_LRELU, dbias=ca.db)
            ca.wgrad(pooled, d1, self.tc, bias_done=True)

    def forward(self, x: torch.Tensor, attn: Optional[List[torch.Tensor]] = None, save: bool = False):
        """x [B,S,S,10] (self.dtype; or already zero-padded to [B,S,S,64] in tensor-core mode) -> y [B,S,S,1].
        attn: per-level features [Ba,...] broadcast as n % Ba."""
        B, S = x.shape[0], x.shape[1]
        if x.shape[-1] == 10:
            want = self.in_channels(B, S, x.shape[2])
            if want == 64:
                x = ops.pad64(x)
            elif want == 16:
                x = ops.pad_channels(x, 16)
        assert x.shape[-1] != 16 or ("enc1a" in self.thin and self.thin["enc1a"].servable(B, S, x.shape[2])), \
            "the 16-channel input form needs the thin first layer to serve this shape (use in_channels(n, h, w) to pick the layout)"
        tape = {"x": x, "enc": [], "dec": [], "bott": []} if save else None
        h, cats = x, []
        for lvl in range(4):
            a, b = self.enc[lvl]
            C = a.conv.cout
            # instance-norm statistics come out of the convolution's epilogue (no separate read of z)
            if lvl == 0 and x.shape[-1] == 16:
                za, sa = self.thin["enc1a"].fwd(h, None, True, self.store.version, want_stats=True)
            else:
                za, sa = self._conv(a.conv, h, stats=True)
            ya, _ = ops.inorm_apply(za, sa, a.gamma, a.beta)
            zb, sb = self._conv(b.conv, ya, stats=True)
            cat = ops.new((B, zb.shape[1], zb.shape[2], 2 * C), self.dtype)
            _, pool = ops.inorm_apply(zb, sb, b.gamma, b.beta, add=None if attn is None else attn[lvl], out=cat[..., C:], pooled=True)
            cats.append(cat)
            if save:
                tape["enc"].append((h, za, sa, ya, zb, sb))
            h = pool
        for bl in self.bott:
            z, sz = self._conv(bl.conv, h, stats=True)
            y, _ = ops.inorm_apply(z, sz, bl.gamma, bl.beta)
            if save:
                tape["bott"].append((h, z, sz))
            h = y
        for u in range(4):
            cat = cats[3 - u]
            C = cat.shape[3] // 2
            self._conv(self.up[u], h, cat[..., :C])
            a, b = self.dec[u]
            za, sa = self._conv(a.conv, cat, stats=True)
            ya, _ = ops.inorm_apply(za, sa, a.gamma, a.beta)
            zb, sb = self._conv(b.conv, ya, stats=True)
            yb, _ = ops.inorm_apply(zb, sb, b.gamma, b.beta)
            if save:
                tape["dec"].append((h, cat, za, sa, ya, zb, sb))
            h = yb
        y = self._conv(self.out, h)
        if save:
            tape["last"] = (h, y)
            return y, tape
        return y

    def _cli_bwd(self, blk: _CLI, x_in, z, sums, dyA=None, dyP=None, need_dx=True):
        dpre = ops.inorm_bwd(z, sums, blk.gamma, dyA, dyP, ACT_LRELU, dbias=blk.conv.db if blk.conv.has_bias else None)
        blk.conv.wgrad(x_in, dpre, self.tc, bias_done=True)
        if not need_dx:
            return None
        return blk.conv.dgrad(dpre, x_in.shape, None, self.tc, self.store.version)

    def grad_range(self, stage):
        """[lo, hi) of the flat gradient buffer that is FINAL once `stage` of the last backward sweep has been enqueued (the buffer is
        in Keras creation order: encoder + attention head, bottleneck, decoder levels 1..4, output layer): ("dec", u) = up{u+1}T ..
        dec{u+1}b (+ the output layer for u = 3), "bott", "head" = everything before the bottleneck."""
        off = self.store.offsets
        if stage == "head":
            return 0, off["bott1.w"][0]
        if stage == "bott":
            return off["bott1.w"][0], off["up1T.w"][0]
        _, u = stage
        return off["up%dT.w" % (u + 1)][0], (off["up%dT.w" % (u + 2)][0] if u < 3 else self.store.n_train)

    def backward(self, tape, dy: torch.Tensor, dattn: Optional[List[torch.Tensor]] = None, attn_nb: int = 0,
                 need_dx: bool = False, hook=None):
        """Accumulates weight gradients into the store; dattn[lvl] (+)= batch-group sums of d(skip) when given.
        Returns d(x) [B,S,S,10] if need_dx.  hook(stage): called when the weight gradients of ("dec", u) / "bott" have been
        enqueued (see grad_range) -- the data-parallel all-reduce of that range starts there."""
        v = self.store.version
        h, y = tape["last"]
        if self.tc and self.out.pw1_ok(h):
            dh = self.out.pw1_bwd(h, dy, y)
        else:
            dpre = ops.act_bwd(dy, y, ACT_LRELU)
            self.out.wgrad(h, dpre, self.tc)
            dh = self.out.dgrad(dpre, h.shape, None, self.tc, v)
        dskips = [None] * 4
        for u in (3, 2, 1, 0):
            hin, cat, za, sa, ya, zb, sb = tape["dec"][u]
            a, b = self.dec[u]
            C = cat.shape[3] // 2
            dya = self._cli_bwd(b, ya, zb, sb, dyA=dh)
            dcat = self._cli_bwd(a, cat, za, sa, dyA=dya)
            dup = ops.act_bwd(dcat[..., :C], cat[..., :C], ACT_LRELU, dbias=self.up[u].db)
            self.up[u].wgrad(hin, dup, self.tc, bias_done=True)
            dh = self.up[u].dgrad(dup, hin.shape, None, self.tc, v)
            dskips[3 - u] = dcat[..., C:]
            if hook is not None:
                hook(("dec", u))
        for i in (1, 0):
            hin, z, sz = tape["bott"][i]
            dh = self._cli_bwd(self.bott[i], hin, z, sz, dyA=dh)
        if hook is not None:
            hook("bott")
        dx = None
        for lvl in (3, 2, 1, 0):
            hin, za, sa, ya, zb, sb = tape["enc"][lvl]
            a, b = self.enc[lvl]
            dya = self._cli_bwd(b, ya, zb, sb, dyA=dskips[lvl], dyP=dh)
            if dattn is not None:
                ops.group_sum(dskips[lvl], attn_nb, dattn[lvl], accumulate=True)
            last = lvl == 0
            dh = self._cli_bwd(a, hin, za, sa, dyA=dya, need_dx=(not last) or need_dx)
            if last:
                dx = dh
        return dx


# ------------------------------------------------------------------------------------------------
# Discriminator
# ------------------------------------------------------------------------------------------------
class Discriminator:
    def __init__(self, image_size, filter_size=64, live_mask=True, dtype=torch.float32, device="cuda", seed=43,
                 tensor_core=True, dropout=0.2):
        self.S, self.live_mask, self.dtype, self.tc, self.dropout = image_size, live_mask, dtype, tensor_core, dropout
        self.pad_in = bool(tensor_core and dtype == torch.bfloat16 and filter_size % 64 == 0)
        self.cin_buf = 64 if self.pad_in else 3
        self.store = ParamStore(discriminator_specs(image_size, filter_size, live_mask), device, seed)
        s, N, cin = self.store, filter_size, 3
        self.blocks = []
        for i, mult in enumerate((1, 2, 4, 8, 16), start=1):
            self.blocks.append(_CLI(s, f"d{i}", cin, N * mult, stride=2, bias=False))
            cin = N * mult
        if live_mask:
            self.attn = (s.bind(Conv("dattn_a", 3, 3, 1, N * 8)), s.bind(Conv("dattn_b", 3, 3, N * 8, N * 8)))
        self.head = s.bind(Conv("head", 3, 3, cin, 1, bias=False))
        self.dense_w = s.views["dense.w"]
        self.dense_dw = s.gviews["dense.w"]
        # tensor-core mode: d1 (3 -> N, 3x3, stride 2) runs as im2col + a 1x1 convolution over the 27 patch values (zero-padded to
        # 64): the Keras kernel (3,3,3,N) IS the [27][N] matrix of that 1x1 layer, so the twin shares the parameter / gradient views
        self.d1_col = None
        if self.pad_in:
            self.d1_col = s.bind(Conv("d1", 1, 1, 27, N, stride=1, act=ACT_LRELU, bias=False))
            s.pad(self.d1_col)
            if live_mask:
                s.pad(self.attn[0])

    def in_channels(self, n, h, w) -> int:
        """The discriminator takes the plain 3-channel image in every mode."""
        return 3

    def _d1_col_ok(self, n, h, w) -> bool:
        return self.d1_col is not None and h % 2 == 0 and w % 2 == 0 and self.d1_col.can_pad(n, h // 2, w // 2)

    def attention(self, mask):
        """MaxPool16(mask) -> 2 x [Conv3x3 + LeakyReLU] @512 (ShmGANwithSSpecSeg.py:358, :404-412)."""
        m = ops.cast(mask, self.dtype) if mask.dtype != self.dtype else mask
        pooled = ops.maxpool(m, 16)
        if self.pad_in and self.attn[0].can_pad(*pooled.shape[:3]):
            pooled = ops.pad64(pooled)
        a1 = self.attn[0].fwd(pooled, None, self.tc, self.store.version)
        a2 = self.attn[1].fwd(a1, None, self.tc, self.store.version)
        return a2, (pooled, a1, a2)

    def attention_backward(self, saved, dattn):
        pooled, a1, a2 = saved
        d2 = ops.act_bwd(dattn, a2, ACT_LRELU, dbias=self.attn[1].db)
        self.attn[1].wgrad(a1, d2, self.tc, bias_done=True)
        d1 = self.attn[1].dgrad(d2, a1.shape, None, self.tc, self.store.version)
        d1 = ops.act_bwd(d1, a1, ACT_LRELU, dbias=self.attn[0].db)
        self.attn[0].wgrad(pooled, d1, self.tc, bias_done=True)

    def forward(self, x, attn=None, noise=None, keep=None, save=False):
        """x [B,S,S,3] -> (rf [B,S/32,S/32,1], cls [B,5] fp32).  noise / keep: the GaussianNoise(0.1) / Dropout(0.2)
        draws of a training=True call (:352, :363); None = training=False."""
        v = self.store.version
        h = x if noise is None else ops.add(x, noise)
        col = self._d1_col_ok(*x.shape[:3])
        if col:
            h = ops.im2col_k3s2(h)                      # [B,S/2,S/2,64]: the 27 patch values of d1, zero-padded
        tape = {"layers": [], "col": col, "hw": (x.shape[1], x.shape[2])} if save else None
        for i, bl in enumerate(self.blocks):
            conv = self.d1_col if (i == 0 and col) else bl.conv
            z, sz = conv.fwd(h, None, self.tc, v, want_stats=True)
            y, _ = ops.inorm_apply(z, sz, bl.gamma, bl.beta, add=attn if i == 3 else None)
            if save:
                tape["layers"].append((h, z, sz))
            h = y
        y5 = h
        if keep is not None:
            h = ops.mul_mask(y5, keep, 1.0 / (1.0 - self.dropout))
        rf = self.head.fwd(h, None, self.tc, v)
        cls = ops.dense_fwd(h, self.dense_w)
        if save:
            tape.update(h5=h, rf=rf, keep=keep)
            return rf, cls, tape
        return rf, cls

    def backward(self, tape, d_rf, d_cls=None, n: Optional[int] = None, wgrad=True, need_dx=False, dattn=None, attn_nb=0):
        """Back-propagates seeds (d_rf [n,s,s,1], d_cls [n,5] fp32 or None) through the first n images of a saved pass.
        wgrad=False gives the dgrad-only sweep that carries the generator loss back to the image."""
        v = self.store.version
        sub = (lambda t: t) if n is None else (lambda t: None if t is None else t[:n])
        h5, rf, keep = sub(tape["h5"]), sub(tape["rf"]), sub(tape["keep"])
        dpre = ops.act_bwd(ops.cast(d_rf, self.dtype), rf, ACT_LRELU)
        if wgrad:
            self.head.wgrad(h5, dpre, self.tc)
        dh = self.head.dgrad(dpre, h5.shape, None, self.tc, v)
        if d_cls is not None:
            if wgrad:
                ops.dense_wgrad(h5, d_cls, self.dense_dw)
            dh = ops.add(dh, ops.dense_dgrad(d_cls, self.dense_w, h5))
        if keep is not None:
            dh = ops.mul_mask(dh, keep, 1.0 / (1.0 - self.dropout))
        for i in (4, 3, 2, 1, 0):
            hin, z, sz = (sub(t) for t in tape["layers"][i])
            bl = self.blocks[i]
            if i == 3 and dattn is not None:
                ops.group_sum(dh, attn_nb, dattn, accumulate=True)
            dpre = ops.inorm_bwd(z, sz, bl.gamma, dyA=dh, act=ACT_LRELU)
            conv = self.d1_col if (i == 0 and tape["col"]) else bl.conv
            if wgrad:
                conv.wgrad(hin, dpre, self.tc)
            if i > 0 or need_dx:
                dh = conv.dgrad(dpre, hin.shape, None, self.tc, v)
        if need_dx and tape["col"]:
            dh = ops.col2im_k3s2(dh, 3, tape["hw"][0], tape["hw"][1], self.dtype)
        return dh if need_dx else None


# ------------------------------------------------------------------------------------------------
# SpecSeg (predict only)
# ------------------------------------------------------------------------------------------------
class SpecSegNet:
    """SpecSeg U-Net (SpecSeg.py:27-98) at predict time: Dropout inactive, BatchNorm on moving statistics.

    In tensor-core mode the 16- and 32-channel levels (encoder levels 1-2, decoder levels 8-9) run in a zero-padded 64-channel
    geometry (ops.PaddedConv): activations are [.., 64] with zeros beyond the real channels, the two decoder concat buffers are
    [.., 128] = [up padded to 64 | skip padded to 64], so every layer of the mask network is a tcgen05 kernel."""

    def __init__(self, dtype=torch.float32, device="cuda", seed=44, tensor_core=True):
        self.dtype, self.tc = dtype, tensor_core
        self.store = ParamStore(specseg_specs(), device, seed, trainable=False)
        s, cin = self.store, 1
        self.enc, self.dec = [], []
        for i, c in enumerate((16, 32, 64, 128, 256), start=1):
            self.enc.append((s.bind(Conv(f"c{i}a", 3, 3, cin, c, act=ACT_RELU)), s.bind(Conv(f"c{i}b", 3, 3, c, c, act=ACT_RELU)), i))
            cin = c
        for i, c in zip((6, 7, 8, 9), (128, 64, 32, 16)):
            self.dec.append((s.bind(Conv(f"u{i}", 2, 2, cin, c, stride=2, transposed=True, act=ACT_NONE)),
                             s.bind(Conv(f"c{i}a", 3, 3, 2 * c, c, act=ACT_RELU)), s.bind(Conv(f"c{i}b", 3, 3, c, c, act=ACT_RELU))))
            cin = c
        self.out = s.bind(Conv("out", 1, 1, 16, 1, act=ACT_SIGMOID))
        # zero-padded twins of the < 64-channel layers (same parameter views)
        self.padded = None
        if tensor_core and dtype == torch.bfloat16:
            P = ops.PaddedConv
            # thin geometry: 16- / 32-channel tensors stay 16 / 32 wide (32- / 64-byte pixel rows in the halo kernel); only the two
            # transposed convs run with 64 padded output columns and store the real ones (nstore) into the [up | skip] concat buffer
            self.padded = {
                "c1a": s.bind(P("c1a", 3, 3, 1, 16, 1, 1, seg_pad=16, cout_dev=16)),
                "c1b": s.bind(P("c1b", 3, 3, 16, 16, 16, 1, seg_pad=16, cout_dev=16)),
                "c2a": s.bind(P("c2a", 3, 3, 16, 32, 16, 1, seg_pad=16, cout_dev=32)),
                "c2b": s.bind(P("c2b", 3, 3, 32, 32, 32, 1, seg_pad=32, cout_dev=32)),
                "c3a": s.bind(P("c3a", 3, 3, 32, 64, 32, 1, seg_pad=32, cout_dev=64)),
                "u8": s.bind(P("u8", 2, 2, 64, 32, 64, 1, stride=2, transposed=True, act=ACT_NONE, nstore=32)),
                "c8a": s.bind(P("c8a", 3, 3, 64, 32, 64, 1, seg_pad=64, cout_dev=32)),
                "c8b": s.bind(P("c8b", 3, 3, 32, 32, 32, 1, seg_pad=32, cout_dev=64)),       # 64 columns (32 zero): u9 reads 64-channel rows
                "u9": s.bind(P("u9", 2, 2, 32, 16, 32, 1, stride=2, transposed=True, act=ACT_NONE, nstore=16)),
                "c9a": s.bind(P("c9a", 3, 3, 32, 16, 32, 1, seg_pad=32, cout_dev=16)),
                "c9b": s.bind(P("c9b", 3, 3, 16, 16, 16, 1, seg_pad=16, cout_dev=16)),
            }
        self._zbuf = {}

    def _zeros(self, key, shape):
        """Zero-initialised scratch whose padding channels are never written (cached per shape across calls)."""
        t = self._zbuf.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = ops.zeros(shape, self.dtype, self.store.flat.device)
            self._zbuf[key] = t
        return t

    def _padded_ok(self, B, H, W) -> bool:
        if self.padded is None or H % 64 != 0 or W % 64 != 0:
            return False
        p = self.padded
        return (p["c1a"].servable(B, H, W) and p["c2a"].servable(B, H // 2, W // 2) and p["c3a"].servable(B, H // 4, W // 4)
                and p["u8"].servable(B, H // 4, W // 4) and p["u9"].servable(B, H // 2, W // 2))

    def predict(self, x: torch.Tensor, verbose=0) -> torch.Tensor:
        """x [B,S,S,1] -> sigmoid probabilities [B,S,S,1] in the same dtype (SpecSeg.predict, ShmGANwithSSpecSeg.py:492)."""
        if self._padded_ok(*x.shape[:3]):
            return self._predict_padded(x)
        v = self.store.version
        sv = self.store.views
        B = x.shape[0]
        h, cats = x, []
        for ca, cb, i in self.enc:
            h = ca.fwd(h, None, self.tc, v)
            h = cb.fwd(h, None, self.tc, v)
            c = cb.cout
            bn = (sv[f"bn{i}.gamma"], sv[f"bn{i}.beta"], sv[f"bn{i}.mean"], sv[f"bn{i}.var"])
            if i < 5:
                cat = ops.new((B, h.shape[1], h.shape[2], 2 * c), self.dtype)     # [up, skip] (SpecSeg.py:65)
                _, h = ops.bn_eval(h, *bn, out=cat[..., c:], pooled=True)
                cats.append(cat)
            else:
                h, _ = ops.bn_eval(h, *bn)
        for (up, ca, cb), cat in zip(self.dec, reversed(cats)):
            c = up.cout
            up.fwd(h, cat[..., :c], self.tc, v)
            h = ca.fwd(cat, None, self.tc, v)
            h = cb.fwd(h, None, self.tc, v)
        return self.out.fwd(h, None, self.tc, v)

    def _predict_padded(self, x: torch.Tensor) -> torch.Tensor:
        v, sv, P = self.store.version, self.store.views, self.padded
        B, H, W = x.shape[:3]
        bn = lambda i: (sv[f"bn{i}.gamma"], sv[f"bn{i}.beta"], sv[f"bn{i}.mean"], sv[f"bn{i}.var"])
        # level 1 (16 channels) and level 2 (32): thin tensors; skips land in the upper half of the [up | skip] concat buffers
        cat9 = ops.new((B, H, W, 32), self.dtype)
        cat8 = ops.new((B, H // 2, W // 2, 64), self.dtype)
        p1 = ops.new((B, H // 2, W // 2, 16), self.dtype)
        p2 = ops.new((B, H // 4, W // 4, 32), self.dtype)
        h = P["c1b"].fwd(P["c1a"].fwd(ops.pad_channels(x, 16), None, True, v), None, True, v)
        ops.bn_eval(h, *bn(1), out=cat9[..., 16:32], pooled=p1)
        h = P["c2b"].fwd(P["c2a"].fwd(p1, None, True, v), None, True, v)
        ops.bn_eval(h, *bn(2), out=cat8[..., 32:64], pooled=p2)
        h = P["c3a"].fwd(p2, None, True, v)
        cats = []
        for ca, cb, i in self.enc[2:]:
            if i > 3:
                h = ca.fwd(h, None, True, v)
            h = cb.fwd(h, None, True, v)
            c = cb.cout
            if i < 5:
                cat = ops.new((B, h.shape[1], h.shape[2], 2 * c), self.dtype)
                _, h = ops.bn_eval(h, *bn(i), out=cat[..., c:], pooled=True)
                cats.append(cat)
            else:
                h, _ = ops.bn_eval(h, *bn(i))
        for (up, ca, cb), cat in zip(self.dec[:2], reversed(cats)):
            c = up.cout
            up.fwd(h, cat[..., :c], True, v)
            h = cb.fwd(ca.fwd(cat, None, True, v), None, True, v)
        P["u8"].fwd(h, cat8[..., :32], True, v)
        h = P["c8b"].fwd(P["c8a"].fwd(cat8, None, True, v), None, True, v)
        P["u9"].fwd(h, cat9[..., :16], True, v)
        h = P["c9b"].fwd(P["c9a"].fwd(cat9, None, True, v), None, True, v)
        return self.out.fwd(h, None, True, v)
