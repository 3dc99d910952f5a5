"""Test-time image-quality metrics of test.py:332-352 on the device: MSE, SSIM, PSNR, mean Delta-E 76 / 94.

`image_metrics(gen_rgb, target_rgb)` returns the row the reference appends to its table per test image.  The kernels are the
training losses' SSIM / min-max kernels (csrc/loss.cu) plus the squared-error and Lab / Delta-E reductions of csrc/extras.cu;
nothing here computes on the host except the final scalar arithmetic on the few reduced values."""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import ops
from ._lib import call
from .ops import _p, _stream


def _minmax3(img):
    n, hw = img.shape[0], img.shape[1] * img.shape[2]
    mm = ops.new((n, 2), torch.float32)
    idx = ops.new((n, 2), torch.int32)
    call("shm_minmax3", _p(img), None, n, hw, _p(mm), _p(idx), _stream())
    return mm


def ssim_rescaled(a: torch.Tensor, b: torch.Tensor, max_val: float = 5.0) -> torch.Tensor:
    """tf.image.ssim(rescale_01(a), rescale_01(b), max_val) -> [N] (test.py:335; rescale_01 utils.py:190-195, per image)."""
    n, h, w, _ = a.shape
    ss = ops.new((n,), torch.float32)
    call("shm_ssim_fwd", _p(a), None, _p(_minmax3(a)), _p(b), _p(_minmax3(b)), n, h, w, float(max_val), _p(ss), None, _stream())
    return ss


def image_metrics(gen_rgb: torch.Tensor, target_rgb: torch.Tensor) -> Dict[str, object]:
    """gen_rgb, target_rgb: [N,H,W,3] fp32 on the device.  Returns per-image lists (and the batch MSE) as Python floats:
    {"mse": float (Keras MeanSquaredError over the whole batch, :343), "ssim": [N], "psnr": [N] (max_val = 1.0, :338),
     "delE76": [N], "delE94": [N] (mean over the pixels of each image, :348-349)}."""
    a, b = gen_rgb.contiguous(), target_rgb.contiguous()
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape and a.shape[3] == 3
    n, h, w, _ = a.shape
    per = h * w * 3
    sq = ops.zeros((n,), torch.float64, a.device)
    call("shm_sqerr_per_image", _p(a), _p(b), n, per, _p(sq), _stream())
    de = ops.zeros((n, 2), torch.float64, a.device)
    call("shm_delta_e", _p(a), _p(b), n, h * w, _p(de), _stream())
    ss = ssim_rescaled(a, b, 5.0)
    sq_h, de_h, ss_h = sq.cpu().tolist(), de.cpu().tolist(), ss.cpu().tolist()
    mse_img = [s / per for s in sq_h]
    return {"mse": sum(sq_h) / (n * per), "ssim": ss_h,
            "psnr": [(-10.0 * math.log10(m)) if m > 0 else float("inf") for m in mse_img],       # 20 log10(1.0) = 0
            "delE76": [d[0] / (h * w) for d in de_h], "delE94": [d[1] / (h * w) for d in de_h]}
