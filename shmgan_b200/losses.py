"""Loss-kernel wrappers (ShmGANwithSSpecSeg.py:669-844).  Every call adds `weight * value` into one slot of a device
scalar table and writes / accumulates the seed gradient scaled by `gscale` (the coefficient of that term in the total
being differentiated).  Values are read back with a single copy at the end of the step."""
from __future__ import annotations

import ctypes as C

import torch

from . import ops
from .ops import _p, _stream, call

SLOTS = ["D1_rf", "D3_rf", "D1_cls", "D3_cls", "D4_cls", "D2_rf", "D4_rf_only", "L1_G1", "L1_c0", "L1_c1", "L1_c2", "L1_c3",
         "L1_c4", "ssim0", "ssim1", "ssim2", "ssim3", "ssim4", "spec0", "spec1", "spec2", "spec3", "spec4", "content", "style"]
IDX = {k: i for i, k in enumerate(SLOTS)}


class LossTable:
    def __init__(self, device="cuda"):
        self.buf = ops.zeros((len(SLOTS),), torch.float32, device)

    def zero(self):
        ops.zero_(self.buf)

    def slot(self, name):
        return C.c_void_p(self.buf.data_ptr() + 4 * IDX[name])

    def read(self):
        host = self.buf.cpu()
        return {k: float(host[i]) for k, i in IDX.items()}


def lsgan(a, target, slot, weight=1.0, da=None, gscale=0.0, accumulate=False):
    """weight * mean((a - target)^2); da (+)= gscale * 2 (a - target) / n.  target: a float, or a one-element fp32 CUDA tensor read when the
    kernel runs (CUDA-graph replays)."""
    if isinstance(target, torch.Tensor):
        call("shm_lsgan_dev", _p(a), a.numel(), _p(target), slot, float(weight), _p(da), float(gscale), int(accumulate), _stream())
    else:
        call("shm_lsgan", _p(a), a.numel(), float(target), slot, float(weight), _p(da), float(gscale), int(accumulate), _stream())


def softmax_ce(logits, labels5, slot, weight=1.0, dlogits=None, gscale=0.0, accumulate=False):
    """labels5: five floats, or a five-element fp32 CUDA tensor read when the kernel runs."""
    if isinstance(labels5, torch.Tensor):
        call("shm_softmax_ce_dev", _p(logits), logits.shape[0], _p(labels5), slot, float(weight), _p(dlogits), float(gscale), int(accumulate), _stream())
        return
    lab = (C.c_float * 5)(*[float(v) for v in labels5])
    call("shm_softmax_ce", _p(logits), logits.shape[0], lab, slot, float(weight), _p(dlogits), float(gscale), int(accumulate), _stream())


def l1(a, b, slot, weight=1.0, da=None, gscale=0.0, accumulate=False):
    call("shm_l1", _p(a), _p(b), a.numel(), slot, float(weight), _p(da), float(gscale), int(accumulate), _stream())


def mse_ycc(Y, cbcr, yuv, slot, weight=1.0, dY=None, gscale=0.0):
    call("shm_mse_ycc", _p(Y), _p(cbcr), _p(yuv), Y.numel(), slot, float(weight), _p(dY), float(gscale), _stream())


def spec(Y, cbcr, yuv, mask, slot, weight=1.0):
    call("shm_spec_loss", _p(Y), _p(cbcr), _p(yuv), _p(mask), Y.numel(), slot, float(weight), _stream())


def minmax3(Y_or_yuv, cbcr):
    n = Y_or_yuv.shape[0]
    hw = Y_or_yuv.shape[1] * Y_or_yuv.shape[2]
    mm = ops.new((n, 2), torch.float32)
    idx = ops.new((n, 2), torch.int32)
    call("shm_minmax3", _p(Y_or_yuv), _p(cbcr), n, hw, _p(mm), _p(idx), _stream())
    return mm, idx


def gram3(Y_or_yuv, cbcr):
    n = Y_or_yuv.shape[0]
    hw = Y_or_yuv.shape[1] * Y_or_yuv.shape[2]
    g = ops.zeros64((n, 9), Y_or_yuv.device)
    call("shm_gram3", _p(Y_or_yuv), _p(cbcr), n, hw, _p(g), _stream())
    return g


def style(Y, cbcr, yuv_ref, S, slot, weight=1.0, dY=None, gscale=0.0):
    """style = 1/(2*9*S^2)^2 * mean((Gram(concat(Y,cbcr)) - Gram(yuv_ref))^2)  (:817-821); dY += gscale * d style / dY."""
    n, h, w, _ = Y.shape
    gA, gB = gram3(Y, cbcr), gram3(yuv_ref, None)
    dg = ops.new((n, 9), torch.float32) if dY is not None else None
    call("shm_style_loss", _p(gA), _p(gB), n, h * w, S, slot, float(weight), _p(dg), float(gscale), _stream())
    if dY is not None:
        call("shm_gram3_bwd", _p(Y), _p(cbcr), n, h * w, _p(dg), _p(dY), _stream())


def ssim_term(Y, cbcr, yuv_ref, slot, weight=1.0, dY=None, gscale=0.0, max_val=5.0):
    """mean_b(-log((1 + ssim_b)/2)) with ssim = tf.image.ssim(rescale_01(concat(Y,cbcr)), rescale_01(yuv_ref), 5)
    (:759-779, utils.py:190-195); dY += gscale * d/dY including the min / max paths of rescale_01."""
    n, h, w, _ = Y.shape
    mmA, idxA = minmax3(Y, cbcr)
    mmB, _ = minmax3(yuv_ref, None)
    ss = ops.new((n,), torch.float32)
    maps = None
    if dY is not None:
        maps = ops.new((int(call("shm_ssim_map_elems", n, h, w)),), torch.float32)
    call("shm_ssim_fwd", _p(Y), _p(cbcr), _p(mmA), _p(yuv_ref), _p(mmB), n, h, w, float(max_val), _p(ss), _p(maps), _stream())
    dss = ops.new((n,), torch.float32) if dY is not None else None
    call("shm_ssim_loss", _p(ss), n, slot, float(weight), _p(dss), float(gscale), _stream())
    if dY is not None:
        scratch = ops.new((n, 2), torch.float64)
        call("shm_ssim_bwd", _p(Y), _p(cbcr), _p(mmA), _p(idxA), _p(yuv_ref), _p(mmB), n, h, w, _p(maps), _p(dss), _p(dY), _p(scratch), _stream())
    return ss
