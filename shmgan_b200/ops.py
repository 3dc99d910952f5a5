"""Thin tensor-level wrappers over the C ABI.  Tensors are torch CUDA tensors used purely as buffer carriers:
NHWC activations, possibly channel-slices of a wider buffer (pixel stride ld = tensor.stride(-2)).
Every function launches libshmgan kernels on torch's current CUDA stream; nothing falls back to torch math."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, BF16, F32, ConvDesc, call  # noqa: F401

IN_EPS = 1e-6       # ShmGANwithSSpecSeg.py:245
BN_EPS = 1e-3       # Keras BatchNormalization default (SpecSeg.py:37)

# When set to a list, every convolution launch is bracketed by CUDA events on the launching stream and appended as
# (family, pass, layer, flops, bytes, ev0, ev1); bench.py uses it for the per-kernel roofline (never on in the timed value).
PROF = None

# Instance-norm statistics from the producing convolution's epilogue (shm_conv2d_tc_fwd_stats) instead of a separate read of the tensor
# (shm_inorm_stats).  Off = the round-1 two-kernel form (kept switchable for A/B measurements and the equivalence test).
import os as _os
FUSE_STATS = _os.environ.get("SHM_FUSE_STATS", "1") != "0"


TC_KERNELS = ("conv_tc_kernel", "conv_halo_kernel", "conv_multi_kernel (big)", "conv_multi_kernel (scatter)", "wgrad_tc_kernel",
              "wgrad_halo_kernel<0>", "wgrad_halo_kernel<1>", "wgrad_s2_kernel<0>", "wgrad_s2_kernel<1>")


def _route(d, pas):
    """Name of the tcgen05 kernel that serves (desc, pass); only evaluated while profiling."""
    if PROF is None:
        return "tc"
    r = call("shm_conv2d_tc_route", C.byref(d), pas)
    return "tc:" + (TC_KERNELS[r] if 0 <= r < len(TC_KERNELS) else "?")


def _prof(family, kind, name, flops, nbytes, fn):
    if PROF is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    PROF.append((family, kind, name, flops, nbytes, e0, e1))
    return r


# fp64 accumulator arena: the ~100 statistics buffers of one step (sum / sum^2 per (n, c), loss partials) are slices of ONE buffer
# that is cleared by a single memset when the step begins, instead of one torch.zeros fill kernel each.  Outside a step
# (arena_begin not called, or the arena is exhausted) zeros64 falls back to torch.zeros.
_ARENA = None
_ARENA_USED = 0
_ARENA_ON = False
ARENA_DOUBLES = 8 << 20


def arena_begin(whole: bool = False):
    """Start of a train / inference step: recycle the arena (clears the part the previous step used, on the current stream).
    whole: clear all of it -- a step being captured into a CUDA graph must not depend on how much the step before it happened to use."""
    global _ARENA, _ARENA_USED, _ARENA_ON
    if _ARENA is None:
        _ARENA = zeros((ARENA_DOUBLES,), torch.float64)
    elif whole:
        zero_(_ARENA)
    elif _ARENA_USED:
        zero_(_ARENA[:_ARENA_USED])
    _ARENA_USED, _ARENA_ON = 0, True


def arena_end():
    global _ARENA_ON
    _ARENA_ON = False


def zeros64(shape, device) -> torch.Tensor:
    global _ARENA_USED
    n = 1
    for v in shape:
        n *= int(v)
    if not _ARENA_ON or _ARENA_USED + n > ARENA_DOUBLES or _ARENA.device != torch.device(device):
        return zeros(shape, torch.float64, device)
    t = _ARENA[_ARENA_USED:_ARENA_USED + n].view(shape)
    _ARENA_USED += (n + 1) // 2 * 2                      # keep 16-byte alignment (double2 loads)
    return t


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_GET_DEV = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  Called once per launch (~600 per step): the raw accessor costs a
    fraction of a microsecond where torch.cuda.current_stream() builds a Stream object (~14 us, most of the host time of a step)."""
    if _RAW_STREAM is None or _GET_DEV is None:
        return torch.cuda.current_stream().cuda_stream
    return _RAW_STREAM(_GET_DEV())


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("unsupported dtype %s" % t.dtype)


def tdtype(code: int):
    return torch.float32 if code == F32 else torch.bfloat16


def ld(t: Optional[torch.Tensor]) -> int:
    """Pixel stride of an NHWC tensor / channel-slice view."""
    if t is None:
        return 0
    assert t.stride(-1) == 1 or t.shape[-1] == 1, "channel axis must be unit-stride"
    return t.stride(-2)


def _check_nhwc(t: torch.Tensor):
    n, h, w, _ = t.shape
    l = t.stride(2)
    assert t.stride(1) == w * l and (n == 1 or t.stride(0) == h * w * l), "tensor is not a dense NHWC pixel grid"


def new(shape, dtype, device="cuda"):
    return torch.empty(shape, dtype=dtype, device=device)


def zero_(t: torch.Tensor) -> torch.Tensor:
    """In-place zero fill of a dense tensor through the library (one memset on the current stream, no framework kernel)."""
    assert t.is_contiguous()
    if t.numel():
        call("shm_zero", _p(t), t.numel() * t.element_size(), _stream())
    return t


def zeros(shape, dtype, device="cuda") -> torch.Tensor:
    return zero_(torch.empty(shape, dtype=dtype, device=device))


def zeros_like(t: torch.Tensor) -> torch.Tensor:
    return zero_(torch.empty(t.shape, dtype=t.dtype, device=t.device))


# ------------------------------------------------------------------------------------------------
# convolution layer: geometry + which kernel family serves it
# ------------------------------------------------------------------------------------------------
class Conv:
    """One Conv2D / Conv2DTranspose layer of the reference (Keras weight layouts), bound to slices of a flat
    parameter / gradient buffer.  `tc` = use the tcgen05 bf16 kernels when the shape allows."""

    def __init__(self, name, kh, kw, cin, cout, stride=1, transposed=False, act=ACT_LRELU, bias=True):
        self.name, self.kh, self.kw, self.cin, self.cout = name, kh, kw, cin, cout
        self.stride, self.transposed, self.act, self.has_bias = stride, transposed, act, bias
        self.w = self.b = self.dw = self.db = None          # views set by ParamStore.bind
        self.w_tc = self.w_tc_d = None                      # bf16 [tap][n][k] copies for the tensor-core path
        self.tc_version = -1
        self.store = None                                   # the ParamStore this layer is bound to (batched weight refresh)
        self.w_tc_d16 = None                                # dgrad layout of a zero-padded first layer at 16 input channels (thin kernel)
        self.tc16_version = -1
        self.cin_pad = None                                 # 64 when the layer reads a zero-padded input on the tensor cores
        self.dw_pad = None                                  # fp32 [kh,kw,cin_pad,cout] weight-gradient scratch of the padded layer

    def enable_pad(self):
        """First layers (Cin = 10 / 3 / 1) in tensor-core mode: the input is presented zero-padded to 64 channels so that the
        reduction dimension fills one 64-channel TMA box; the padded weight rows are zero, so the arithmetic is unchanged."""
        assert not self.transposed and self.cin < 64 and self.cout % 64 == 0
        self.cin_pad = 64
        return self

    def can_pad(self, n, h, w) -> bool:
        """True when the zero-padded tensor-core form can serve an [n,h,w,*] input (the lattice must tile into 128-point boxes)."""
        if self.cin_pad is None:
            return False
        d = self.desc(n, h, w, self.cin_pad, self.cout, BF16, cin=self.cin_pad)
        return self.tc_ok(d)

    def padded(self, channels: int) -> bool:
        return self.cin_pad is not None and channels == self.cin_pad and self.cin != self.cin_pad

    def pw1_ok(self, x) -> bool:
        """1x1 conv to one channel on bf16 activations: served by the bandwidth kernel (shm_pw1_*)."""
        return (self.kh == 1 and self.kw == 1 and self.cout == 1 and not self.transposed and x.dtype == torch.bfloat16
                and self.cin in (8, 16, 32, 64, 128, 256) and ld(x) % 8 == 0 and x.data_ptr() % 16 == 0)

    def c3to1_ok(self, x) -> bool:
        """3x3 stride-1 conv to one channel on bf16 activations (the discriminator's real/fake head): warp-per-pixel dot products."""
        return (self.kh == 3 and self.kw == 3 and self.cout == 1 and self.stride == 1 and not self.transposed
                and x.dtype == torch.bfloat16 and self.cin % 8 == 0 and x.shape[-1] == self.cin and ld(x) % 8 == 0
                and x.data_ptr() % 16 == 0)

    def wshape(self):
        return (self.kh, self.kw, self.cout, self.cin) if self.transposed else (self.kh, self.kw, self.cin, self.cout)

    def out_hw(self, h, w):
        if self.transposed:
            return h * self.stride, w * self.stride
        return -(-h // self.stride), -(-w // self.stride)

    def desc(self, n, h, w, ldx, ldy, dtype, act=None, cin=None) -> ConvDesc:
        return ConvDesc(n, h, w, self.cin if cin is None else cin, self.cout, self.kh, self.kw, self.stride, int(self.transposed),
                        self.act if act is None else act, ldx, ldy, dtype, 0)

    # -- tensor-core weights ---------------------------------------------------------------------
    def refresh_tc(self, version: int):
        """Re-lays the fp32 master weights out as bf16 [tap][n][k] for fwd and dgrad (once per optimiser step)."""
        if self.tc_version == version:
            return
        cin = self.cin_pad or self.cin
        d = self.desc(1, 16, 16, cin, self.cout, BF16, cin=cin)
        if self.w_tc is None:
            n = self.kh * self.kw * cin * self.cout
            self.w_tc = new((n,), torch.bfloat16)
            self.w_tc_d = new((n,), torch.bfloat16)
            if self.store is not None:
                self.store.register_tc(self)                # from the next optimiser step on, refreshed by the store's one batched launch
        call("shm_conv2d_tc_prep_weights_both", C.byref(d), _p(self.w), self.cin, _p(self.w_tc), _p(self.w_tc_d), _stream())
        self.tc_version = version

    def prep_job(self) -> bytes:
        """This layer's record for the batched weight refresh (shm_conv2d_tc_prep_multi)."""
        cin = self.cin_pad or self.cin
        d = self.desc(1, 16, 16, cin, self.cout, BF16, cin=cin)
        buf = C.create_string_buffer(int(call("shm_conv2d_tc_prep_job_bytes")))
        call("shm_conv2d_tc_prep_job", C.byref(d), _p(self.w), self.cin, _p(self.w_tc), _p(self.w_tc_d), buf)
        return buf.raw

    def tc_ok(self, d: ConvDesc) -> bool:
        return (d.dtype == BF16 and d.Cin % 64 == 0 and self.cout % 64 == 0
                and bool(call("shm_conv2d_tc_supported", C.byref(d), 0)))

    # -- forward / backward ------------------------------------------------------------------------
    def fwd(self, x: torch.Tensor, y: Optional[torch.Tensor] = None, tc: bool = True, version: int = 0, want_stats: bool = False):
        """want_stats: also return the instance-norm statistics of y ([n, cout, 2] fp64 sums / sums of squares per image and channel) --
        from the convolution's own epilogue when the serving kernel has one, else from a separate pass over y."""
        n, h, w, _ = x.shape
        ho, wo = self.out_hw(h, w)
        if y is None:
            y = new((n, ho, wo, self.cout), x.dtype)
        pad = self.padded(x.shape[-1])
        d = self.desc(n, h, w, ld(x), ld(y), dt(x), cin=self.cin_pad if pad else None)
        fl, nb = self.flops(n, h, w), self.io_bytes(n, h, w, x.element_size())
        if pad and not (tc and self.tc_ok(d)):
            raise L.ShmError("%s: zero-padded input needs the tensor-core path (shape %s)" % (self.name, tuple(x.shape)))
        if want_stats:
            sums = None
            if tc and FUSE_STATS and not self.c3to1_ok(x) and not self.pw1_ok(x) and self.tc_ok(d) and call("shm_conv2d_tc_stats_supported", C.byref(d)):
                self.refresh_tc(version)
                sums = zeros64((n, self.cout, 2), x.device)
                _prof(_route(d, 0), "fwd", self.name, fl, nb, lambda: call("shm_conv2d_tc_fwd_stats", C.byref(d), _p(x), _p(self.w_tc), _p(self.b), _p(y), _p(sums), _stream()))
            else:
                self.fwd(x, y, tc, version)
                sums = inorm_stats(y)
            return y, sums
        if tc and self.c3to1_ok(x):
            _prof("bw", "fwd", self.name, fl, nb, lambda: call("shm_c3to1_fwd", _p(x), n, h, w, self.cin, ld(x), _p(self.w), _p(self.b), self.act,
                                                              _p(y), dt(x), _stream()))
        elif tc and self.pw1_ok(x):
            _prof("bw", "fwd", self.name, fl, nb, lambda: call("shm_pw1_fwd", _p(x), ld(x), self.cin, _p(self.w), _p(self.b), self.act, _p(y),
                                                              n * h * w, dt(x), _stream()))
        elif tc and self.tc_ok(d):
            self.refresh_tc(version)
            _prof(_route(d, 0), "fwd", self.name, fl, nb, lambda: call("shm_conv2d_tc_fwd", C.byref(d), _p(x), _p(self.w_tc), _p(self.b), _p(y), _stream()))
        else:
            _prof("simt", "fwd", self.name, fl, nb, lambda: call("shm_conv2d_fwd", C.byref(d), _p(x), _p(self.w), _p(self.b), _p(y), _stream()))
        return y

    def flops(self, n, h, w):
        """Algorithmic FLOPs (2 x MACs) of one pass over an [n,h,w,cin] input (same count for fwd, dgrad and wgrad)."""
        if self.transposed:
            return 2 * n * h * w * self.cin * self.cout * self.kh * self.kw
        ho, wo = self.out_hw(h, w)
        return 2 * n * ho * wo * self.cin * self.cout * self.kh * self.kw

    def io_bytes(self, n, h, w, esize):
        """Algorithmic HBM bytes of one pass: read the input once, write the output once (weights are negligible)."""
        ho, wo = self.out_hw(h, w)
        return esize * n * (h * w * self.cin + ho * wo * self.cout)

    def thin_dgrad(self, dy: torch.Tensor, x_shape, version: int):
        """dgrad of a zero-padded first layer through the thin halo kernel: dx comes out 16 channels wide (the real ones first)
        instead of cin_pad = 64 -- a quarter of the MMAs and of the bytes.  Returns None when the shape is not servable."""
        n, h, w, _ = x_shape
        if self.cin > 16 or self.transposed or self.stride != 1 or self.kh != 3 or self.cout % 64 != 0:
            return None
        d = ConvDesc(n, h, w, 16, self.cout, 3, 3, 1, 0, ACT_NONE, 16, ld(dy), BF16, 0)
        if dy.dtype != torch.bfloat16 or not call("shm_conv2d_tc_supported", C.byref(d), 1):
            return None
        if self.w_tc_d16 is None:
            self.w_tc_d16 = new((9 * 16 * self.cout,), torch.bfloat16)
            self.tc16_version = -1
        if self.tc16_version != version:
            call("shm_conv2d_tc_prep_weights", C.byref(d), _p(self.w), self.cin, _p(self.w_tc_d16), 1, _stream())
            self.tc16_version = version
        dx = new((n, h, w, 16), dy.dtype)
        fl, nb = self.flops(n, h, w), self.io_bytes(n, h, w, dy.element_size())
        _prof(_route(d, 1), "dgrad", self.name, fl, nb, lambda: call("shm_conv2d_tc_dgrad", C.byref(d), _p(dy), _p(self.w_tc_d16), _p(dx), _stream()))
        return dx

    def dgrad(self, dy: torch.Tensor, x_shape, dx: Optional[torch.Tensor] = None, tc: bool = True, version: int = 0):
        """dx from dy = dL/d(pre-activation)."""
        n, h, w, cx = x_shape
        pad = self.padded(cx)
        if (pad or (self.cin_pad is not None and cx == 16)) and tc and dx is None:
            thin = self.thin_dgrad(dy, x_shape, version)
            if thin is not None:
                return thin
            if not pad:
                raise L.ShmError("%s: the 16-channel input form needs the thin dgrad kernel" % self.name)
        if dx is None:
            dx = new((n, h, w, cx if pad else self.cin), dy.dtype)
        d = self.desc(n, h, w, ld(dx), ld(dy), dt(dy), ACT_NONE, cin=self.cin_pad if pad else None)
        fl, nb = self.flops(n, h, w), self.io_bytes(n, h, w, dy.element_size())
        if pad and not (tc and self.tc_ok(d)):
            raise L.ShmError("%s: zero-padded dgrad needs the tensor-core path" % self.name)
        if tc and self.c3to1_ok(dx):
            _prof("bw", "dgrad", self.name, fl, nb, lambda: call("shm_c3to1_dgrad", _p(dy), n, h, w, self.cin, _p(self.w), _p(dx), ld(dx), dt(dy), _stream()))
        elif tc and self.tc_ok(d):
            self.refresh_tc(version)
            _prof(_route(d, 1), "dgrad", self.name, fl, nb, lambda: call("shm_conv2d_tc_dgrad", C.byref(d), _p(dy), _p(self.w_tc_d), _p(dx), _stream()))
        else:
            _prof("simt", "dgrad", self.name, fl, nb, lambda: call("shm_conv2d_dgrad", C.byref(d), _p(dy), _p(self.w), _p(dx), 0, _stream()))
        return dx

    def wgrad(self, x: torch.Tensor, dy: torch.Tensor, tc: bool = True, bias_done: bool = False):
        """dw += , db += (gradients accumulate: weights are shared by several passes).
        bias_done: db was already accumulated by the kernel that produced dy (inorm_bwd / act_bwd with dbias=)."""
        n, h, w, cx = x.shape
        # a zero-padded first layer may be given its input 16 channels wide (dense): the halo wgrad kernel's 64-channel TMA box then
        # runs past the channel extent and is zero-filled on the way into shared memory (ldx = 16 < Cin = 64 in the descriptor)
        pad = self.padded(cx) or (self.cin_pad is not None and cx == 16 and x.is_contiguous())
        d = self.desc(n, h, w, ld(x), ld(dy), dt(x), ACT_NONE, cin=self.cin_pad if pad else None)
        fl, nb = self.flops(n, h, w), self.io_bytes(n, h, w, x.element_size())
        if pad and not (tc and self.tc_ok(d)):
            raise L.ShmError("%s: zero-padded wgrad needs the tensor-core path" % self.name)
        dw = self.dw
        if pad:
            if self.dw_pad is None:
                self.dw_pad = zeros((self.kh, self.kw, self.cin_pad, self.cout), torch.float32, x.device)
            dw = self.dw_pad
        if tc and self.c3to1_ok(x) and not self.has_bias:
            _prof("bw", "wgrad", self.name, fl, nb, lambda: call("shm_c3to1_wgrad", _p(x), n, h, w, self.cin, ld(x), _p(dy), _p(dw), dt(x), _stream()))
        elif tc and self.tc_ok(d):
            _prof(_route(d, 2), "wgrad", self.name, fl, nb, lambda: call("shm_conv2d_tc_wgrad", C.byref(d), _p(x), _p(dy), _p(dw), _stream()))
            if self.has_bias and not bias_done:
                call("shm_colsum", _p(dy), dy.shape[0] * dy.shape[1] * dy.shape[2], self.cout, ld(dy), dt(dy), _p(self.db), _stream())
        else:
            _prof("simt", "wgrad", self.name, fl, nb, lambda: call("shm_conv2d_wgrad", C.byref(d), _p(x), _p(dy), _p(self.dw), _p(self.db) if (self.has_bias and not bias_done) else None, _stream()))

    def pw1_bwd(self, x, dy, y, need_dx=True):
        """Fused backward of the 1x1 -> 1 channel layer: activation derivative, dx, dw +=, db += in one pass."""
        n, h, w, _ = x.shape
        dx = new((n, h, w, self.cin), x.dtype) if need_dx else None
        fl, nb = 2 * self.flops(n, h, w), 2 * self.io_bytes(n, h, w, x.element_size())
        _prof("bw", "bwd", self.name, fl, nb, lambda: call("shm_pw1_bwd", _p(x), ld(x), self.cin, _p(self.w), _p(dy), _p(y), self.act, _p(dx), ld(dx),
                                                          _p(self.dw), _p(self.db) if self.has_bias else None, n * h * w, dt(x), _stream()))
        return dx

    def zero_pad_grad(self):
        if self.dw_pad is not None:
            zero_(self.dw_pad)

    def fold_pad_grad(self):
        """dw[tap, :cin, :] = dw_pad[tap, :cin, :] (the padded rows are gradients of weights that do not exist)."""
        if self.dw_pad is None:
            return
        taps, n = self.kh * self.kw, self.cin * self.cout
        call("shm_cast2d", _p(self.dw_pad), F32, self.cin_pad * self.cout, _p(self.dw), F32, n, taps, n, _stream())


class PaddedConv(Conv):
    """Inference-only layer whose channel counts are below the tensor-core granule (SpecSeg's 16/32-channel levels,
    SpecSeg.py:34-44, :76-86): it runs in a zero-padded device geometry -- inputs live in 64-channel segments holding
    `seg_real` real channels each, outputs are `cout_dev` wide with zero weights / bias beyond `cout` -- so that the 3x3 layers
    go through the halo tcgen05 kernel instead of the exact-fp32 SIMT kernel (6.9 ms -> < 1 ms per step for the mask network).
    The padded output channels are exactly zero after ReLU / LeakyReLU / no activation (zero weights, zero bias)."""

    def __init__(self, name, kh, kw, cin, cout, seg_real, nseg, stride=1, transposed=False, act=ACT_RELU, bias=True,
                 seg_pad=64, cout_dev=None, nstore=None):
        """seg_pad / cout_dev below 64 select the THIN halo kernel (32- / 64-byte pixel rows, 16 / 32 output columns) for stride-1
        3x3 layers; nstore (transposed layers) = output channels actually stored (the rest of the cout_dev columns are padding)."""
        super().__init__(name, kh, kw, cin, cout, stride=stride, transposed=transposed, act=act, bias=bias)
        assert seg_real * nseg == cin and act != ACT_SIGMOID and seg_pad >= seg_real
        self.seg_real, self.nseg, self.seg_pad = seg_real, nseg, seg_pad
        self.cin_dev, self.cout_dev = seg_pad * nseg, (max(64, cout) if cout_dev is None else cout_dev)
        assert self.cout_dev >= cout
        self.nstore = nstore
        self.b_dev = None

    def refresh_tc(self, version: int):
        if self.tc_version == version:
            return
        d = ConvDesc(1, 16, 16, self.cin_dev, self.cout_dev, self.kh, self.kw, self.stride, int(self.transposed), self.act,
                     self.cin_dev, self.cout_dev, BF16, 0)
        if self.w_tc is None:
            self.w_tc = new((self.kh * self.kw * self.cin_dev * self.cout_dev,), torch.bfloat16)
            self.b_dev = zeros((self.cout_dev,), torch.float32, self.w.device)
        call("shm_conv2d_tc_prep_weights_padded", C.byref(d), _p(self.w), self.cin, self.seg_real, self.seg_pad, self.cout, _p(self.w_tc), _stream())
        if self.has_bias:
            call("shm_cast", _p(self.b), F32, _p(self.b_dev), F32, self.cout, _stream())
        self.tc_version = version

    def dev_desc(self, n, h, w, ldx, ldy):
        return ConvDesc(n, h, w, self.cin_dev, self.cout_dev, self.kh, self.kw, self.stride, int(self.transposed), self.act, ldx, ldy, BF16, 0)

    def servable(self, n, h, w) -> bool:
        return bool(call("shm_conv2d_tc_supported", C.byref(self.dev_desc(n, h, w, self.cin_dev, self.cout_dev)), 0))

    def fwd(self, x, y=None, tc=True, version=0, want_stats=False):
        """x [n,h,w,cin_dev] bf16 (zero-padded segments) -> y [n,ho,wo,cout_dev] bf16 (channels cout.. are zero)."""
        n, h, w, cx = x.shape
        assert tc and x.dtype == torch.bfloat16 and cx == self.cin_dev, (self.name, tuple(x.shape))
        ho, wo = self.out_hw(h, w)
        nst = self.nstore or self.cout_dev
        if y is None:
            y = new((n, ho, wo, nst), x.dtype)
        assert y.shape[-1] == nst
        self.refresh_tc(version)
        d = self.dev_desc(n, h, w, ld(x), ld(y))
        fl, nb = self.flops(n, h, w), x.element_size() * n * (h * w * self.cin_dev + ho * wo * nst)
        if want_stats:
            assert nst == self.cout_dev == self.cout, "statistics are for layers whose device columns are all real channels"
            if FUSE_STATS and call("shm_conv2d_tc_stats_supported", C.byref(d)):
                sums = zeros64((n, self.cout, 2), x.device)
                _prof(_route(d, 0), "fwd", self.name, fl, nb, lambda: call("shm_conv2d_tc_fwd_stats", C.byref(d), _p(x), _p(self.w_tc), _p(self.b_dev), _p(y), _p(sums), _stream()))
                return y, sums
            self.fwd(x, y, tc, version)
            return y, inorm_stats(y)
        if nst != self.cout_dev:
            _prof(_route(d, 0), "fwd", self.name, fl, nb, lambda: call("shm_conv2d_tc_fwd_cols", C.byref(d), _p(x), _p(self.w_tc), _p(self.b_dev), _p(y), nst, _stream()))
        else:
            _prof(_route(d, 0), "fwd", self.name, fl, nb, lambda: call("shm_conv2d_tc_fwd", C.byref(d), _p(x), _p(self.w_tc), _p(self.b_dev), _p(y), _stream()))
        return y


# ------------------------------------------------------------------------------------------------
# instance norm (+pool, +add, +slice placement) and friends
# ------------------------------------------------------------------------------------------------
def inorm_stats(x: torch.Tensor) -> torch.Tensor:
    n, h, w, c = x.shape
    sums = zeros64((n, c, 2), x.device)
    _prof("norm", "stats", "c%d" % c, 0, x.numel() * x.element_size(),
          lambda: call("shm_inorm_stats", _p(x), n, h * w, c, ld(x), dt(x), _p(sums), _stream()))
    return sums


def inorm_apply(x, sums, gamma, beta, add=None, out=None, pooled=False, want_out=True):
    """Returns (out, pooled).  add: [nadd,H,W,C] broadcast over the batch as n % nadd."""
    n, h, w, c = x.shape
    if out is None and want_out:
        out = new((n, h, w, c), x.dtype)
    pl = new((n, h // 2, w // 2, c), x.dtype) if pooled else None
    nb = x.numel() * x.element_size() * ((1 if out is None else 2) + (0 if add is None else 1) + (0.25 if pooled else 0))
    _prof("norm", "apply", "c%d" % c, 0, nb,
          lambda: call("shm_inorm_apply", _p(x), n, h, w, c, ld(x), dt(x), _p(sums), _p(gamma), _p(beta), IN_EPS,
                       _p(add), ld(add), 0 if add is None else add.shape[0], _p(out), ld(out), _p(pl), ld(pl), _stream()))
    return out, pl


def inorm_bwd(x, sums, gamma, dyA=None, dyP=None, act=ACT_LRELU, dx=None, dbias=None):
    """dL/d(pre-activation of the producing conv) from dL/dy, y = IN(x), x = post-activation conv output.
    dbias (fp32 [C], optional) += column sums of the result: the producing conv's bias gradient, fused into the same pass."""
    n, h, w, c = x.shape
    bs = zeros64((n, c, 2), x.device)
    e = x.numel() * x.element_size()
    rd = e * (1 + (0 if dyA is None else 1) + (0 if dyP is None else 0.25))
    _prof("norm", "bwd_stats", "c%d" % c, 0, rd,
          lambda: call("shm_inorm_bwd_stats", _p(x), n, h, w, c, ld(x), dt(x), _p(sums), IN_EPS, _p(dyA), ld(dyA), _p(dyP), ld(dyP), _p(bs), _stream()))
    if dx is None:
        dx = new((n, h, w, c), x.dtype)
    _prof("norm", "bwd_apply", "c%d" % c, 0, rd + e,
          lambda: call("shm_inorm_bwd_apply", _p(x), n, h, w, c, ld(x), dt(x), _p(sums), _p(gamma), IN_EPS, _p(dyA), ld(dyA), _p(dyP), ld(dyP),
                       _p(bs), act, _p(dx), ld(dx), _p(dbias), _stream()))
    return dx


def act_bwd(dy, y, act, out=None, dbias=None):
    """dpre = dy * act'(y); dbias (fp32 [C], optional) += column sums of dpre (the conv's bias gradient, same pass)."""
    n, h, w, c = y.shape
    if out is None:
        out = new((n, h, w, c), y.dtype)
    _prof("norm", "act_bwd", "c%d" % c, 0, 3 * y.numel() * y.element_size(),
          lambda: call("shm_act_bwd", _p(dy), ld(dy), _p(y), ld(y), _p(out), ld(out), n * h * w, c, act, dt(y), _p(dbias), _stream()))
    return out


def maxpool(x, k):
    n, h, w, c = x.shape
    y = new((n, h // k, w // k, c), x.dtype)
    call("shm_maxpool", _p(x), n, h, w, c, ld(x), k, _p(y), ld(y), dt(x), _stream())
    return y


def bn_eval(x, gamma, beta, mean, var, out=None, pooled=False):
    """pooled: False | True (allocate) | a [n,h/2,w/2,c] tensor / channel-slice view to write MaxPool2 of the result into."""
    n, h, w, c = x.shape
    if out is None:
        out = new((n, h, w, c), x.dtype)
    if torch.is_tensor(pooled):
        pl = pooled
    else:
        pl = new((n, h // 2, w // 2, c), x.dtype) if pooled else None
    call("shm_bn_eval", _p(x), n, h, w, c, ld(x), dt(x), _p(gamma), _p(beta), _p(mean), _p(var), BN_EPS, _p(out), ld(out), _p(pl), ld(pl), _stream())
    return out, pl


def add(a, b, out=None):
    n, h, w, c = a.shape
    if out is None:
        out = new((n, h, w, c), a.dtype)
    call("shm_add", _p(a), ld(a), _p(b), ld(b), _p(out), ld(out), n * h * w, c, dt(a), _stream())
    return out


def group_sum(src, nb, dst, accumulate):
    """dst[b] (+)= sum_r src[r*nb + b]."""
    n, h, w, c = src.shape
    call("shm_group_sum", _p(src), ld(src), n // nb, nb * h * w, c, _p(dst), ld(dst), int(accumulate), dt(src), _stream())
    return dst


def mul_mask(x, keep, scale):
    out = torch.empty_like(x)
    call("shm_mul_mask", _p(x), _p(keep), _p(out), x.numel(), float(scale), dt(x), _stream())
    return out


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    assert x.is_contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    call("shm_cast", _p(x), dt(x), _p(out), dt(out), x.numel(), _stream())
    return out


def cast_into(src, dst):
    """dst[...] = src[...] for NHWC tensors / channel-slice views of equal shape (strided copy + dtype conversion)."""
    c = src.shape[-1]
    call("shm_cast2d", _p(src), dt(src), ld(src), _p(dst), dt(dst), ld(dst), src.numel() // c, c, _stream())
    return dst


def axpy(alpha, x, y):
    """y += alpha * x (dense tensors of one dtype)."""
    assert x.is_contiguous() and y.is_contiguous() and x.dtype == y.dtype and x.numel() == y.numel()
    call("shm_axpy", float(alpha), _p(x), _p(y), x.numel(), dt(x), _stream())
    return y


def rng_normal(shape, seed, offset, sigma, dtype):
    """offset: an int, or a one-element int64 CUDA tensor read when the kernel runs (CUDA-graph replays)."""
    out = new(shape, dtype)
    if isinstance(offset, torch.Tensor):
        call("shm_rng_normal_dev", _p(out), out.numel(), seed, _p(offset), float(sigma), dt(out), _stream())
    else:
        call("shm_rng_normal", _p(out), out.numel(), seed, offset, float(sigma), dt(out), _stream())
    return out


def rng_keep(shape, seed, offset, keep_prob, dtype):
    out = new(shape, dtype)
    if isinstance(offset, torch.Tensor):
        call("shm_rng_keep_dev", _p(out), out.numel(), seed, _p(offset), float(keep_prob), dt(out), _stream())
    else:
        call("shm_rng_keep", _p(out), out.numel(), seed, offset, float(keep_prob), dt(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# dense head
# ------------------------------------------------------------------------------------------------
def dense_fwd(x, w):
    b = x.shape[0]
    k = x.numel() // b
    out = new((b, w.shape[1]), torch.float32)
    call("shm_dense_fwd", _p(x), _p(w), _p(out), b, k, w.shape[1], dt(x), _stream())
    return out


def dense_dgrad(dout, w, like):
    dx = torch.empty_like(like)
    b = like.shape[0]
    call("shm_dense_dgrad", _p(dout), _p(w), _p(dx), b, like.numel() // b, w.shape[1], dt(dx), _stream())
    return dx


def dense_wgrad(x, dout, dw):
    b = x.shape[0]
    call("shm_dense_wgrad", _p(x), _p(dout), _p(dw), b, x.numel() // b, dw.shape[1], dt(x), _stream())


# ------------------------------------------------------------------------------------------------
# preprocessing
# ------------------------------------------------------------------------------------------------
def pseudo_diffuse_min4(i0, i45, i90, i135):
    """calculate_estimate_diffuse (utils.py:102-106).  fp32 / bf16 / uint8."""
    code = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2}[i0.dtype]
    out = torch.empty_like(i0)
    if i0.numel() == 0:                      # empty input -> empty output (numpy semantics of the reference), no launch
        return out
    call("shm_pseudo_diffuse_min4", _p(i0), _p(i45), _p(i90), _p(i135), _p(out), i0.numel(), code, _stream())
    return out


def yuv_standardize(rgb, scale_out=None):
    """rgb [N,H,W,3] fp32 -> (standardised yuv [N,H,W,3] fp32, scale [N]); scale_out: a dense [N] fp32 view to write the scales into."""
    n, h, w, _ = rgb.shape
    sums = zeros64((n, 2), rgb.device)
    call("shm_yuv_stats", _p(rgb), n, h * w, _p(sums), _stream())
    yuv = torch.empty_like(rgb)
    scale = new((n,), torch.float32) if scale_out is None else scale_out
    assert scale.numel() == n and scale.is_contiguous() and scale.dtype == torch.float32
    call("shm_yuv_standardize", _p(rgb), n, h * w, _p(sums), _p(yuv), _p(scale), _stream())
    return yuv, scale


class RunningMean:
    """Device-side stand-in for the reference's unbounded `self.stddev_arr` list (ShmGANwithSSpecSeg.py:1306, datasetLoader.py:42): `append`
    adds a tensor of scales to an fp64 (sum, count) pair on the device; `scaled(x, mul)` = x * mul * mean without a host round trip
    (tf.reduce_mean(self.stddev_arr) at :548 / test.py:246); `mean()` reads it back (synchronises)."""

    def __init__(self, device="cuda"):
        self.acc = zeros((2,), torch.float64, device)

    def append(self, t: torch.Tensor):
        assert t.dtype == torch.float32 and t.is_contiguous()
        call("shm_sum_count", _p(t), t.numel(), _p(self.acc), _stream())

    def scaled(self, x: torch.Tensor, mul: float) -> torch.Tensor:
        assert x.dtype == torch.float32 and x.is_contiguous()
        out = torch.empty_like(x)
        call("shm_scale_by_mean", _p(x), _p(self.acc), float(mul), _p(out), x.numel(), _stream())
        return out

    def mean(self) -> float:
        a = self.acc.cpu()
        return float(a[0] / a[1]) if float(a[1]) > 0 else 0.0

    def __len__(self) -> int:
        return int(self.acc[1].item())

    def state(self):
        return self.acc.cpu().tolist()

    def load_state(self, st):
        self.acc.copy_(torch.tensor(list(st), dtype=torch.float64))


def avg_cbcr(ds: Sequence[torch.Tensor]):
    n, h, w, _ = ds[0].shape
    out = new((n, h, w, 2), torch.float32)
    call("shm_avg_cbcr", _p(ds[0]), _p(ds[1]), _p(ds[2]), _p(ds[3]), _p(ds[4]), _p(out), n * h * w, _stream())
    return out


def assemble_input(srcs, src_lds, onehot, out):
    """srcs: 5 fp32 tensors (or None = zeros) giving one value per pixel at stride src_lds[j]; out [N,H,W,10] or the
    zero-padded [N,H,W,64] tensor-core form."""
    arr_p = (C.c_void_p * 5)(*[None if s is None else s.data_ptr() for s in srcs])
    arr_l = (C.c_int32 * 5)(*src_lds)
    ldo = out.shape[-1]
    call("shm_assemble_input", C.cast(arr_p, C.POINTER(C.c_void_p)), arr_l, onehot, _p(out), ldo, out.numel() // ldo, dt(out), _stream())
    return out


def assemble_bwd(din, slots, dgen):
    arr = (C.c_int32 * 5)(*(list(slots) + [0] * (5 - len(slots))))
    ldin = din.shape[-1]
    call("shm_assemble_bwd", _p(din), dt(din), ldin, arr, len(slots), _p(dgen), din.numel() // ldin, _stream())


def pad64(src, out=None):
    """[N,H,W,C] (C <= 64, fp32 or bf16, any pixel stride) -> dense bf16 [N,H,W,64] with channels C.. zero."""
    n, h, w, c = src.shape
    if out is None:
        out = new((n, h, w, 64), torch.bfloat16)
    assert out.is_contiguous() and out.shape[-1] == 64 and out.dtype == torch.bfloat16
    call("shm_pad_channels64", _p(src), dt(src), ld(src), c, _p(out), n * h * w, _stream())
    return out


def pad_channels(src, cpad: int, out=None):
    """[N,H,W,C] (C <= cpad in {16, 32, 64}) -> dense bf16 [N,H,W,cpad] with channels C.. zero (inputs of the thin tensor-core layers)."""
    n, h, w, c = src.shape
    if out is None:
        out = new((n, h, w, cpad), torch.bfloat16)
    assert out.is_contiguous() and out.shape[-1] == cpad and out.dtype == torch.bfloat16
    call("shm_pad_channels", _p(src), dt(src), ld(src), c, _p(out), cpad, n * h * w, _stream())
    return out


def im2col_k3s2(x):
    """[N,H,W,C] (C <= 7, fp32 / bf16, any pixel stride) -> bf16 [N,H/2,W/2,64] patch tensor of a 3x3 stride-2 SAME conv."""
    n, h, w, c = x.shape
    out = new((n, h // 2, w // 2, 64), torch.bfloat16)
    call("shm_im2col_k3s2", _p(x), dt(x), ld(x), n, h, w, c, _p(out), _stream())
    return out


def col2im_k3s2(dP, c, h, w, dtype=torch.bfloat16):
    """Transpose of im2col_k3s2: bf16 [N,H/2,W/2,64] -> [N,H,W,c]."""
    n = dP.shape[0]
    assert dP.is_contiguous() and dP.shape[-1] == 64 and dP.dtype == torch.bfloat16
    dx = new((n, h, w, c), dtype)
    call("shm_col2im_k3s2", _p(dP), n, h, w, c, _p(dx), dt(dx), ld(dx), _stream())
    return dx


def add_channels_(a, b):
    """a[..., :C] += b  in place (C = b's channel count; a may be wider, e.g. the zero-padded discriminator input)."""
    c = b.shape[-1]
    call("shm_add", _p(a), ld(a), _p(b), ld(b), _p(a), ld(a), b.numel() // c, c, dt(a), _stream())
    return a


def yuv2rgb(Y, cbcr, rgb=None, lp=None):
    """Y [N,H,W,1] fp32, cbcr [Nc,H,W,2] fp32 broadcast over N as n % Nc -> rgb [N,H,W,3] fp32 and / or a bf16 copy `lp`
    (the discriminator's input in bf16 mode).  Outputs must be dense."""
    n, h, w, _ = Y.shape
    call("shm_yuv2rgb", _p(Y), _p(cbcr), cbcr.numel() // 2, _p(rgb), _p(lp), BF16 if lp is not None else F32,
         3 if lp is None else lp.shape[-1], n * h * w, _stream())
    return rgb, lp


def yuv2rgb_bwd(drgb_f32, drgb_lp, dY, accumulate):
    code = dt(drgb_lp) if drgb_lp is not None else F32
    call("shm_yuv2rgb_bwd", _p(drgb_f32), _p(drgb_lp), code, 3 if drgb_lp is None else drgb_lp.shape[-1], _p(dY), dY.numel(),
         int(accumulate), _stream())


# ------------------------------------------------------------------------------------------------
# rows SURVEY 8(f): loader contract, degree of polarisation (csrc/extras.cu)
# ------------------------------------------------------------------------------------------------
def load_u8_images(img_u8: torch.Tensor, image_size: int, flip_ud: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """datasetLoader.py:48-62 on the device: uint8 [N,Hs,Ws,3] -> bilinear resize (TF2 half-pixel) -> /255 -> optional vertical flip."""
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[3] == 3, "uint8 [N,H,W,3] expected"
    n, hs, ws, _ = img_u8.shape
    if out is None:
        out = new((n, image_size, image_size, 3), torch.float32)
    if n == 0:                                # empty batch -> empty output, no launch
        return out
    call("shm_load_u8_bilinear", _p(img_u8.contiguous()), n, hs, ws, _p(out), image_size, image_size, int(bool(flip_ud)), _stream())
    return out


def dop(i0, i45, i90, i135, want_angle: bool = False):
    """calcDOP (ShmGANwithSSpecSeg.py:1157-1169) on fp32 planes of any shape; returns DoP (and the angle of polarisation)."""
    out = torch.empty_like(i0)
    aop = torch.empty_like(i0) if want_angle else None
    call("shm_dop", _p(i0.contiguous()), _p(i45.contiguous()), _p(i90.contiguous()), _p(i135.contiguous()), _p(out), _p(aop), i0.numel(), _stream())
    return (out, aop) if want_angle else out
