"""CPU restatement (numpy) of the rows SURVEY.md section 8(f) marks "next": the dataset-loader contract, the test-time
image-quality metrics and the degree of polarisation.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED against TensorFlow (not installable here); pinned instead
against independent implementations that ARE in this image (tests/test_extras_cpu.py): torch's half-pixel bilinear interpolation
for the resize, OpenCV's float RGB->Lab for the colour conversion, and hand-computed values for the Delta-E formulas.

Every function cites the reference lines it follows (paths relative to the reference checkout, Atif-Anwer/SHMGAN) and, where the
arithmetic lives in a third-party dependency, the published algorithm it restates.
"""
from __future__ import annotations

import numpy as np

__all__ = ["resize_bilinear_tf2", "load_images", "dop", "rgb_to_lab", "delta_e76", "delta_e94", "psnr", "mse", "image_metrics"]


# --------------------------------------------------------------------------------------
# dataset loader contract (datasetLoader.py:48-62)
# --------------------------------------------------------------------------------------
def _interp_weights(out_size: int, in_size: int):
    """TF 2.8 `compute_interpolation_weights` with HalfPixelScaler (tf.image.resize(method='bilinear'), the resize behind
    keras.preprocessing.image_dataset_from_directory(image_size=...), datasetLoader.py:48-57): all in float32."""
    scale = np.float32(in_size) / np.float32(out_size)
    o = np.arange(out_size, dtype=np.float32)
    src = (o + np.float32(0.5)) * scale - np.float32(0.5)
    lo_f = np.floor(src)
    lower = np.maximum(lo_f, 0).astype(np.int64)
    upper = np.minimum(np.ceil(src), in_size - 1).astype(np.int64)
    lerp = (src - lo_f).astype(np.float32)
    return lower, upper, lerp


def resize_bilinear_tf2(img_u8: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [N,Hs,Ws,C] -> float32 [N,out_h,out_w,C]; TF's ResizeBilinear kernel order of operations:
    top = tl + (tr - tl) * xl; bottom = bl + (br - bl) * xl; out = top + (bottom - top) * yl, each a separate float32 rounding."""
    x = img_u8.astype(np.float32)
    ylo, yhi, yl = _interp_weights(out_h, x.shape[1])
    xlo, xhi, xl = _interp_weights(out_w, x.shape[2])
    xl = xl[None, None, :, None]
    yl = yl[None, :, None, None]
    tl, tr = x[:, ylo][:, :, xlo], x[:, ylo][:, :, xhi]
    bl, br = x[:, yhi][:, :, xlo], x[:, yhi][:, :, xhi]
    top = (tl + ((tr - tl) * xl).astype(np.float32)).astype(np.float32)
    bot = (bl + ((br - bl) * xl).astype(np.float32)).astype(np.float32)
    return (top + ((bot - top) * yl).astype(np.float32)).astype(np.float32)


def load_images(img_u8: np.ndarray, image_size: int, random_flip: bool) -> np.ndarray:
    """datasetLoader.py:48-62: resize to (image_size, image_size), x / 255.0 (:60), and `x if random_flip else flip_up_down(x)`
    (:61 -- the flip happens when random_flip is FALSE)."""
    x = resize_bilinear_tf2(img_u8, image_size, image_size) / np.float32(255.0)
    return x if random_flip else x[:, ::-1]


# --------------------------------------------------------------------------------------
# degree of polarisation (ShmGANwithSSpecSeg.py:1157-1169)
# --------------------------------------------------------------------------------------
def dop(i0, i45, i90, i135):
    """calcDOP: S0 = I0 + I90, S1 = I0 - I90, S2 = I45 - I135; DoP = divide_no_nan(sqrt(S1^2 + S2^2), S0) (:1158-1163);
    the angle 0.5 * atan2(S2, S1) is evaluated and discarded by the reference (:1164) -- returned here as the second value."""
    s0, s1, s2 = i0 + i90, i0 - i90, i45 - i135
    pol = np.sqrt(s1 * s1 + s2 * s2)
    d = np.where(s0 == 0, np.zeros_like(pol), pol / np.where(s0 == 0, np.ones_like(s0), s0))
    return d.astype(i0.dtype), (0.5 * np.arctan2(s2, s1)).astype(i0.dtype)


# --------------------------------------------------------------------------------------
# test-time metrics (test.py:332-352)
# --------------------------------------------------------------------------------------
_XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
_D65 = np.array([0.95047, 1.0, 1.08883])


def rgb_to_lab(rgb):
    """tfio.experimental.color.rgb_to_lab (test.py:346-347; tensorflow-io, unpinned): sRGB -> XYZ (inverse gamma, the CIE RGB matrix
    above) -> Lab with the D65 / 2-degree white point; the same published algorithm as skimage.color.rgb2lab."""
    v = np.asarray(rgb, dtype=np.float64)
    lin = np.where(v > 0.04045, ((v + 0.055) / 1.055) ** 2.4, v / 12.92)
    xyz = lin @ _XYZ_FROM_RGB.T / _D65
    f = np.where(xyz > 0.008856, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    L = 116.0 * f[..., 1] - 16.0
    a = 500.0 * (f[..., 0] - f[..., 1])
    b = 200.0 * (f[..., 1] - f[..., 2])
    return np.stack([L, a, b], axis=-1)


def delta_e76(lab1, lab2):
    """skimage.color.deltaE_cie76 (imported by test.py, called :348): Euclidean distance in Lab, per pixel."""
    d = np.asarray(lab1, np.float64) - np.asarray(lab2, np.float64)
    return np.sqrt((d * d).sum(axis=-1))


def delta_e94(lab1, lab2, kH=1.0, kC=1.0, kL=1.0, k1=0.045, k2=0.015):
    """skimage.color.deltaE_ciede94 with its graphic-arts defaults (test.py:349): the first colour is the reference."""
    l1, a1, b1 = [np.asarray(lab1, np.float64)[..., i] for i in range(3)]
    l2, a2, b2 = [np.asarray(lab2, np.float64)[..., i] for i in range(3)]
    c1, c2 = np.hypot(a1, b1), np.hypot(a2, b2)
    dL, dC = l1 - l2, c1 - c2
    dH2 = 2.0 * (c1 * c2 - (a1 * a2 + b1 * b2))
    sc, sh = 1.0 + k1 * c1, 1.0 + k2 * c1
    de2 = (dL / kL) ** 2 + (dC / (kC * sc)) ** 2 + dH2 / (kH * sh) ** 2
    return np.sqrt(np.maximum(de2, 0.0))


def mse(a, b):
    """tf.keras.losses.MeanSquaredError()(a, b) (test.py:342-343): mean over every element."""
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return float((d * d).mean())


def psnr(a, b, max_val=1.0):
    """tf.image.psnr (test.py:338): per image, 20 log10(max_val) - 10 log10(mean squared error over H, W, C) -> [B]."""
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    m = (d * d).reshape(d.shape[0], -1).mean(axis=1)
    return 20.0 * np.log10(max_val) - 10.0 * np.log10(m)


def image_metrics(gen_rgb, target_rgb):
    """The per-image metric row of test.py:332-352: MSE, SSIM (of the globally rescaled images, max_val = 5), PSNR, mean dE76, mean dE94."""
    import torch
    from . import shmgan_oracle as O
    g, t = torch.from_numpy(np.asarray(gen_rgb, np.float64)), torch.from_numpy(np.asarray(target_rgb, np.float64))
    s = O.ssim(O.rescale_01(g, per_image=False), O.rescale_01(t, per_image=False), 5.0).numpy()
    l1, l2 = rgb_to_lab(gen_rgb), rgb_to_lab(target_rgb)
    return {"mse": mse(gen_rgb, target_rgb), "ssim": s, "psnr": psnr(gen_rgb, target_rgb, 1.0),
            "delE76": float(delta_e76(l1, l2).mean()), "delE94": float(delta_e94(l1, l2).mean())}
