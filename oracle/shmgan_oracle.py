"""CPU restatement (PyTorch, float64 or float32) of the SHMGAN hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED (no TensorFlow here).

Every function cites the reference lines it follows (paths relative to the reference
checkout, Atif-Anwer/SHMGAN).  Layout conventions are the reference's: activations NHWC,
Conv2D kernels (kh, kw, Cin, Cout), Conv2DTranspose kernels (kh, kw, Cout, Cin), Dense
(in, out) with Flatten in H, W, C order.  Internally tensors are permuted to NCHW only to
call torch's conv primitives.

TensorFlow 2.8 semantics restated here:
  * SAME padding: out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0);
    before = pad_total // 2 (the odd element goes at the END).
  * Conv2DTranspose(k, s=2, SAME) is the input-gradient of that SAME conv: the full
    transposed convolution cropped at the end to 2*in.
  * tfa InstanceNormalization(axis=-1, epsilon=1e-6): per (n, c) biased moments over H*W.
  * tf.nn.leaky_relu default alpha = 0.2.
  * Keras Adam: theta -= lr_t * m / (sqrt(v) + eps), lr_t = lr * sqrt(1-b2^t) / (1-b1^t).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "tf_same_pad", "conv2d_same", "conv2d_transpose_same", "instance_norm", "leaky_relu",
    "avg_pool2", "max_pool", "generator_param_specs", "discriminator_param_specs",
    "specseg_param_specs", "init_params", "generator_forward", "attention_features",
    "discriminator_forward", "specseg_forward", "rgb_to_yuv", "yuv_to_rgb",
    "per_image_standardization", "rescale_01", "ssim", "gram_matrix", "pseudo_diffuse_min4",
    "assemble_g1_input", "assemble_cyclic_inputs", "train_step_losses", "softmax_ce", "train_step_grads",
    "keras_adam_lr", "keras_adam_update", "inference_step", "count_params",
    "RGB2YUV", "YUV2RGB", "LRELU_ALPHA", "IN_EPS", "BN_EPS", "bf16_storage", "bf16_storage_bwd",
]

LRELU_ALPHA = 0.2      # tf.nn.leaky_relu default (ShmGANwithSSpecSeg.py:244 activation=tf.nn.leaky_relu)
IN_EPS = 1e-6          # ShmGANwithSSpecSeg.py:245 epsilon=0.000001
BN_EPS = 1e-3          # Keras BatchNormalization default (SpecSeg.py:37)

# tf.image.rgb_to_yuv / yuv_to_rgb kernels (ShmGANwithSSpecSeg.py:480, :553); out = x @ K
RGB2YUV = [[0.299, -0.14714119, 0.61497538],
           [0.587, -0.28886916, -0.51496512],
           [0.114, 0.43601035, -0.10001026]]
YUV2RGB = [[1.0, 1.0, 1.0],
           [0.0, -0.394642334, 2.03206185],
           [1.13988303, -0.58062185, 0.0]]


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------
def tf_same_pad(size: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME' rule -> (out, pad_before, pad_after)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    before = total // 2
    return out, before, total - before


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv2d_same(x, w, b=None, stride: int = 1):
    """Keras Conv2D(padding='same').  x NHWC, w (kh,kw,Cin,Cout).  ShmGANwithSSpecSeg.py:244,387."""
    kh, kw = w.shape[0], w.shape[1]
    _, pt, pb = tf_same_pad(x.shape[1], kh, stride)
    _, pl, pr = tf_same_pad(x.shape[2], kw, stride)
    xp = F.pad(_nchw(x), (pl, pr, pt, pb))
    y = F.conv2d(xp, w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    return _nhwc(y)


def conv2d_transpose_same(x, w, b=None, stride: int = 2):
    """Keras Conv2DTranspose(padding='same', strides=2).  w (kh,kw,Cout,Cin).

    ShmGANwithSSpecSeg.py:298 (k=3) and SpecSeg.py:64 (k=2).  out[p] = sum_{o,k: s*o+k-pad_before=p} ...
    where pad_before is the SAME pad_before of the matching forward conv on the 2x-size image
    (0 for k=3,s=2 and k=2,s=2) => full transposed conv cropped at the end.
    """
    kh, kw = w.shape[0], w.shape[1]
    H, W = x.shape[1], x.shape[2]
    _, pt, _ = tf_same_pad(H * stride, kh, stride)
    _, pl, _ = tf_same_pad(W * stride, kw, stride)
    y = F.conv_transpose2d(_nchw(x), w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    y = y[:, :, pt:pt + H * stride, pl:pl + W * stride]
    return _nhwc(y)


def leaky_relu(x, alpha: float = LRELU_ALPHA):
    return torch.where(x > 0, x, x * alpha)


def instance_norm(x, gamma, beta, eps: float = IN_EPS):
    """tfa.layers.InstanceNormalization(axis=-1).  Decomposition: Generator_summary.txt:9-37."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def avg_pool2(x):
    """AveragePooling2D(2,2,'same') on even sizes.  ShmGANwithSSpecSeg.py:249."""
    return _nhwc(F.avg_pool2d(_nchw(x), 2))


def max_pool(x, k: int = 2):
    """MaxPooling2D(k, strides=None -> k, 'same') on sizes divisible by k.  :406."""
    return _nhwc(F.max_pool2d(_nchw(x), k))


# --------------------------------------------------------------------------------------
# parameter inventories (Keras creation order)
# --------------------------------------------------------------------------------------
def generator_param_specs(filter_size: int = 64, live_mask: bool = True):
    """[(name, shape, kind)] in Keras layer-creation order.  ShmGANwithSSpecSeg.py:228-327, :404-412.

    kind: 'w' conv kernel N(0,.02) (:200), 'b' bias zeros, 'g' IN gamma ones, 'be' IN beta N(0,.02)
    (frozen buffers).  With live_mask=False the attention convs are absent, which is the
    as-written graph (Generator_summary.txt has no conv2d_2,3,6,7,10,11,14,15).
    """
    specs = []
    N = filter_size
    cin = 10
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


def discriminator_param_specs(image_size: int, filter_size: int = 64, live_mask: bool = True):
    """ShmGANwithSSpecSeg.py:343-389.  No biases anywhere in D (use_bias=False)."""
    specs = []
    N = filter_size
    cin = 3
    for i, mult in enumerate((1, 2, 4, 8), start=1):
        specs += [(f"d{i}.w", (3, 3, cin, N * mult), "w"),
                  (f"d{i}.in_gamma", (N * mult,), "g"), (f"d{i}.in_beta", (N * mult,), "be")]
        cin = N * mult
    if live_mask:
        specs += [("dattn_a.w", (3, 3, 1, cin), "w"), ("dattn_a.b", (cin,), "b"),
                  ("dattn_b.w", (3, 3, cin, cin), "w"), ("dattn_b.b", (cin,), "b")]
    specs += [("d5.w", (3, 3, cin, N * 16), "w"),
              ("d5.in_gamma", (N * 16,), "g"), ("d5.in_beta", (N * 16,), "be")]
    cin = N * 16
    specs += [("head.w", (3, 3, cin, 1), "w")]
    s32 = image_size // 32
    specs += [("dense.w", (s32 * s32 * cin, 5), "w")]
    return specs


def specseg_param_specs():
    """SpecSeg.py:27-98.  kinds: 'w5' N(0,.05) ('RandomNormal' string), 'glorot', BN 'g','z','m','v'."""
    specs = []
    ch = [16, 32, 64, 128, 256]
    cin = 1
    for i, c in enumerate(ch, start=1):
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


def init_params(specs, seed: int = 42, dtype=torch.float64, randomize_all: bool = False):
    """Seeded synthetic parameters.  randomize_all=True also perturbs biases / BN stats so that parity
    tests exercise every term (the reference's initial biases are zero)."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape, kind in specs:
        if kind == "w":
            t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.02
        elif kind == "w5":
            t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.05
        elif kind == "glorot":
            rf = 1
            for d in shape[:-2]:
                rf *= d
            fan_in, fan_out = shape[-1] * rf, shape[-2] * rf      # (kh,kw,Cout,Cin) for ConvT
            if len(shape) == 4 and shape[0] == 1:                 # 1x1 Conv2D (kh,kw,Cin,Cout)
                fan_in, fan_out = shape[-2], shape[-1]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        elif kind == "be":
            t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.02
        elif kind == "g":
            t = torch.ones(shape, dtype=torch.float64)
            if randomize_all and name.startswith("bn"):          # BN gamma / moving variance
                t = t + 0.3 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif kind in ("b", "z"):
            t = torch.zeros(shape, dtype=torch.float64)
            if randomize_all:
                t = torch.randn(shape, generator=g, dtype=torch.float64) * 0.05
        else:
            raise ValueError(kind)
        out[name] = t.to(dtype)
    return out


def count_params(specs, trainable_only: bool = False) -> int:
    n = 0
    for name, shape, kind in specs:
        if kind in ("g", "be") and ("in_gamma" in name or "in_beta" in name):
            continue  # tfa IN affine is untracked in the reference summaries
        if trainable_only and (name.endswith(".mean") or name.endswith(".var")):
            continue
        k = 1
        for d in shape:
            k *= d
        n += k
    return n


# --------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------
def _ident(t):
    return t


def bf16_storage(t):
    """Storage-precision model of the bf16 execution mode: round to the bf16 grid where the CUDA path stores an
    activation, with a straight-through gradient.  Passed as `q=` to the model functions below it gives the reference the
    bf16 kernels are held to (same quantisation points, exact arithmetic in between); `q=None` is the plain reference."""
    r = t.detach().to(torch.float32).to(torch.bfloat16).to(t.dtype)
    return t + (r - t.detach())


class _Bf16BothWays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.float32).to(torch.bfloat16).to(g.dtype)


def bf16_storage_bwd(t):
    """bf16_storage that ALSO rounds the gradient flowing back through the same point to the bf16 grid: the CUDA path stores the
    backward tensors (d/d pre-activation, d/d layer input) in bf16 as well, at the same places it stores activations.  This is the
    storage model the whole-network bf16 gradient bounds are derived from (tests/test_gpu_nets.py, tests/test_gpu_baseline_shapes.py)."""
    return _Bf16BothWays.apply(t)


def _cli_parts(p, name, x, stride=1, bias=True, q=_ident):
    """Conv -> (+bias) -> LeakyReLU -> InstanceNorm (ShmGANwithSSpecSeg.py:244-245 / :386-389); returns the un-stored
    normalised tensor (the caller stores it, possibly after a fused add / pool)."""
    z = q(leaky_relu(conv2d_same(x, p[name + ".w"], p[name + ".b"] if bias else None, stride)))
    return instance_norm(z, p[name + ".in_gamma"], p[name + ".in_beta"])


def _cli(p, name, x, stride=1, bias=True, q=_ident):
    return q(_cli_parts(p, name, x, stride, bias, q))


def attention_features(p, mask, prefix="attn", levels=(1, 2, 3, 4), q=None):
    """attention_layer (:404-412) evaluated on a live mask: level 1 un-pooled, then MaxPool2 chain."""
    q = q or _ident
    feats = []
    pooled = mask
    for lvl in levels:
        if lvl > 1:
            pooled = max_pool(pooled, 2)
        a = q(leaky_relu(conv2d_same(pooled, p[f"{prefix}{lvl}a.w"], p[f"{prefix}{lvl}a.b"])))
        a = q(leaky_relu(conv2d_same(a, p[f"{prefix}{lvl}b.w"], p[f"{prefix}{lvl}b.b"])))
        feats.append(a)
    return feats


def generator_forward(p, x, mask=None, return_intermediates: bool = False, q=None):
    """build_generator (:228-327).  mask=None reproduces the as-written graph (attn == 0, Q1).
    The attention features are computed first and added where each skip is formed, which is the same arithmetic as the
    reference's later `down_k + attn_k` (:290-293)."""
    q = q or _ident
    inter = {}
    attn = attention_features(p, mask, q=q) if mask is not None else None
    if attn is not None:
        inter["attn"] = attn
    skips = []
    h = x
    for lvl in range(1, 5):
        h = _cli(p, f"enc{lvl}a", h, q=q)
        y = _cli_parts(p, f"enc{lvl}b", h, q=q)
        skips.append(q(y + attn[lvl - 1]) if attn is not None else q(y))      # :290-293 (add, not multiply)
        h = q(avg_pool2(y))                                                    # :249
    h = _cli(p, "bott1", h, q=q)
    h = _cli(p, "bott2", h, q=q)
    for u in range(1, 5):
        up = q(leaky_relu(conv2d_transpose_same(h, p[f"up{u}T.w"], p[f"up{u}T.b"], 2)))   # :298
        h = torch.cat([up, skips[4 - u]], dim=3)                                        # :299
        h = _cli(p, f"dec{u}a", h, q=q)
        h = _cli(p, f"dec{u}b", h, q=q)
    y = q(leaky_relu(conv2d_same(h, p["out.w"], p["out.b"])))                           # :326
    if return_intermediates:
        inter["skips"] = skips
        return y, inter
    return y


def discriminator_forward(p, x, mask=None, training: bool = False, noise=None, keep=None,
                          dropout_rate: float = 0.2, q=None):
    """build_discriminator (:343-380).  noise: N(0,0.1) tensor added to x when training (:352);
    keep: {0,1} tensor for Dropout(0.2) on the d5 output when training (:363), scaled 1/(1-rate)."""
    q = q or _ident
    h = x
    if training and noise is not None:
        h = q(h + noise)
    for i in (1, 2, 3):
        h = _cli(p, f"d{i}", h, stride=2, bias=False, q=q)
    y4 = _cli_parts(p, "d4", h, stride=2, bias=False, q=q)
    if mask is not None:
        pooled = max_pool(mask, 16)                                                     # :358
        a = q(leaky_relu(conv2d_same(pooled, p["dattn_a.w"], p["dattn_a.b"])))
        a = q(leaky_relu(conv2d_same(a, p["dattn_b.w"], p["dattn_b.b"])))
        y4 = y4 + a                                                                     # :359
    h = q(y4)
    h = _cli(p, "d5", h, stride=2, bias=False, q=q)
    if training and keep is not None:
        h = q(h * keep * (1.0 / (1.0 - dropout_rate)))
    rf = q(leaky_relu(conv2d_same(h, p["head.w"], None, 1)))                            # :365-369
    cls = h.reshape(h.shape[0], -1) @ p["dense.w"]                                      # :371-375
    return rf, cls


def specseg_forward(p, x, q=None):
    """SpecSeg U-Net at predict time (SpecSeg.py:27-98): dropout inactive, BN uses moving stats."""
    q = q or _ident

    def bn(h, i):
        return (h - p[f"bn{i}.mean"]) * torch.rsqrt(p[f"bn{i}.var"] + BN_EPS) * p[f"bn{i}.gamma"] + p[f"bn{i}.beta"]
    skips = []
    h = x
    for i in range(1, 6):
        h = q(F.relu(conv2d_same(h, p[f"c{i}a.w"], p[f"c{i}a.b"])))
        h = q(F.relu(conv2d_same(h, p[f"c{i}b.w"], p[f"c{i}b.b"])))
        h = q(bn(h, i))
        if i < 5:
            skips.append(h)
            h = max_pool(h, 2)
    for i in (6, 7, 8, 9):
        up = q(conv2d_transpose_same(h, p[f"u{i}.w"], p[f"u{i}.b"], 2))
        h = torch.cat([up, skips[9 - i]], dim=3)
        h = q(F.relu(conv2d_same(h, p[f"c{i}a.w"], p[f"c{i}a.b"])))
        h = q(F.relu(conv2d_same(h, p[f"c{i}b.w"], p[f"c{i}b.b"])))
    return q(torch.sigmoid(conv2d_same(h, p["out.w"], p["out.b"])))


# --------------------------------------------------------------------------------------
# colour / preprocessing / losses
# --------------------------------------------------------------------------------------
def rgb_to_yuv(x):
    return x @ torch.tensor(RGB2YUV, dtype=x.dtype)


def yuv_to_rgb(x):
    return x @ torch.tensor(YUV2RGB, dtype=x.dtype)


def per_image_standardization(x, per_image: bool = True):
    """custom_per_image_standardization (:1271-1309): x / max(std, rsqrt(65536)); NO mean subtraction.
    per_image=True takes the stats per sample (identical at the reference's batch 1)."""
    dims = (1, 2, 3) if per_image else (0, 1, 2, 3)
    mean = x.mean(dim=dims, keepdim=True)
    var = F.relu((x * x).mean(dim=dims, keepdim=True) - mean * mean)
    scale = torch.clamp(torch.sqrt(var), min=1.0 / math.sqrt(65536.0))
    return x / scale, scale


def rescale_01(x, per_image: bool = True):
    """utils.py:190-195 (divide_no_nan)."""
    dims = (1, 2, 3) if per_image else (0, 1, 2, 3)
    mn = x.amin(dim=dims, keepdim=True)
    mx = x.amax(dim=dims, keepdim=True)
    d = mx - mn
    return torch.where(d != 0, (x - mn) / torch.where(d != 0, d, torch.ones_like(d)), torch.zeros_like(x))


def _gauss_kernel(size=11, sigma=1.5, dtype=torch.float64):
    coords = torch.arange(size, dtype=torch.float64) - (size - 1) / 2.0
    g = -0.5 * coords * coords / (sigma * sigma)
    g2 = (g[None, :] + g[:, None]).reshape(-1)
    return torch.softmax(g2, dim=0).reshape(size, size).to(dtype)


def ssim(img1, img2, max_val: float, filter_size=11, sigma=1.5, k1=0.01, k2=0.03):
    """tf.image.ssim (TF 2.8 image_ops_impl._ssim_per_channel).  NHWC -> [B].  Call site :759."""
    C = img1.shape[3]
    k = _gauss_kernel(filter_size, sigma, img1.dtype)[None, None].repeat(C, 1, 1, 1)

    def red(t):
        return F.conv2d(_nchw(t), k, groups=C)

    c1 = (k1 * max_val) ** 2
    c2 = (k2 * max_val) ** 2
    m0, m1 = red(img1), red(img2)
    num0 = m0 * m1 * 2.0
    den0 = m0 * m0 + m1 * m1
    lum = (num0 + c1) / (den0 + c1)
    num1 = red(img1 * img2) * 2.0
    den1 = red(img1 * img1 + img2 * img2)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (lum * cs).mean(dim=(2, 3)).mean(dim=1)


def gram_matrix(x):
    """:1176-1180."""
    return torch.einsum("bijc,bijd->bcd", x, x) / float(x.shape[1] * x.shape[2])


def pseudo_diffuse_min4(i0, i45, i90, i135):
    """calculate_estimate_diffuse (utils.py:102-106): per pixel, per colour channel min of the 4."""
    return torch.minimum(torch.minimum(i0, i45), torch.minimum(i90, i135))


def softmax_ce(labels, logits):
    return -(labels * torch.log_softmax(logits, dim=1)).sum(dim=1)


# --------------------------------------------------------------------------------------
# train step
# --------------------------------------------------------------------------------------
def assemble_g1_input(Y: Sequence[torch.Tensor], bits: Sequence[bool]):
    """:509-531.  Y = 5 x [B,S,S,1]; bit set -> slot zeroed; one-hot plane = ED (0,0,0,0,1)."""
    z = torch.zeros_like(Y[0])
    o = torch.ones_like(Y[0])
    planes = [z if bits[k] else Y[k] for k in range(5)]
    return torch.cat(planes + [z, z, z, z, o], dim=3)


def assemble_cyclic_inputs(Y, gen_Y, bits):
    """:576-594.  Slot k zeroed; dropped slots carry gen_Y; one-hot plane k."""
    z = torch.zeros_like(Y[0])
    o = torch.ones_like(Y[0])
    sub = [gen_Y if bits[k] else Y[k] for k in range(5)]
    outs = []
    for k in range(5):
        planes = [z if j == k else sub[j] for j in range(5)]
        onehot = [o if j == k else z for j in range(5)]
        outs.append(torch.cat(planes + onehot, dim=3))
    return outs


def train_step_losses(Gp, Dp, origs: Sequence[torch.Tensor], mask, bits: Sequence[bool], T: float,
                      d_noise: Optional[Sequence[torch.Tensor]] = None,
                      d_keep: Optional[Sequence[torch.Tensor]] = None,
                      live_mask: bool = True, per_image: bool = True, q=None):
    """The taped region of train_step (:495-844).  origs = (orig0, orig45, orig90, orig135, origED),
    each [B,S,S,3] in [0,1].  mask = SpecSeg.predict(I90_Ych) (:492), passed in (outside the tape).
    d_noise / d_keep: the GaussianNoise / Dropout draws for the two training=True D calls (D1, D2).
    Batch semantics: CE / SSIM terms are means over the batch.
    q: storage-precision model (`bf16_storage`) applied where the bf16 execution mode stores network inputs / activations."""
    q = q or _ident
    S = origs[0].shape[1]
    gmask = mask if live_mask else None
    ds = [per_image_standardization(rgb_to_yuv(o), per_image)[0] for o in origs]        # :480-484
    Y = [d[..., 0:1] for d in ds]                                                       # :486-490
    avgCbCr = (ds[0][..., 1:] + ds[1][..., 1:] + ds[2][..., 1:] + ds[3][..., 1:] + ds[4][..., 1:]) / 5.0
    out = {"ds_yuv": ds, "avgCbCr": avgCbCr}

    gen_input = q(assemble_g1_input(Y, bits))                                           # :531
    gen_Y = generator_forward(Gp, gen_input, gmask, q=q)                                # :538
    gen_rgb = yuv_to_rgb(torch.cat([gen_Y, avgCbCr], dim=3))                            # :544,553
    n1 = d_noise[0] if d_noise is not None else None
    n2 = d_noise[1] if d_noise is not None else None
    k1 = d_keep[0] if d_keep is not None else None
    k2 = d_keep[1] if d_keep is not None else None
    rf_gen, cls_gen = discriminator_forward(Dp, q(gen_rgb), gmask, True, n1, k1, q=q)   # :559
    rf_tgt, cls_tgt = discriminator_forward(Dp, q(origs[4]), gmask, True, n2, k2, q=q)  # :563

    cyc_in = [q(c) for c in assemble_cyclic_inputs(Y, gen_Y, bits)]                     # :576-594
    cyc_Y = [generator_forward(Gp, ci, gmask, q=q) for ci in cyc_in]                    # :603-607
    cyc_yuv = [torch.cat([cy, avgCbCr], dim=3) for cy in cyc_Y]                         # :613-617
    cyc_rgb = [yuv_to_rgb(c) for c in cyc_yuv]                                          # :620-624
    d3 = [discriminator_forward(Dp, q(c), gmask, False, q=q) for c in cyc_rgb]          # :627-631
    d4 = [discriminator_forward(Dp, q(o), gmask, False, q=q) for o in origs]            # :638-642

    def sqd(a, t):
        return ((a - t) ** 2).mean()

    D3_rf = sum(sqd(r, T) for r, _ in d3)                                               # :669-674
    D1_rf = sqd(rf_gen, T)                                                              # :677
    G_gan = (D3_rf + D1_rf) / 6.0                                                       # :679
    eye = torch.eye(5, dtype=origs[0].dtype)
    D3_cls = sum(softmax_ce(eye[k][None], c).mean() for k, (_, c) in enumerate(d3))    # :695-700
    tgt_lbl = torch.zeros(1, 5, dtype=origs[0].dtype)
    tgt_lbl[0, 4] = T                                                                   # :477,:688
    D1_cls = softmax_ce(tgt_lbl, cls_gen).mean()                                        # :702
    G_clsf = (D3_cls + D1_cls) / 6.0                                                    # :704
    D4_cls = sum(softmax_ce(eye[k][None], c).mean() for k, (_, c) in enumerate(d4))    # :709-714
    D2_rf = sqd(rf_tgt, T) + (rf_gen ** 2).mean()                                       # :721
    D4_rf = sum(sqd(d4[k][0], T) + (d3[k][0] ** 2).mean() for k in range(5)) + D2_rf    # :723-728

    L1_G1 = (gen_rgb - origs[4]).abs().mean()                                           # :744
    L1_c = [(cyc_rgb[k] - origs[k]).abs().mean() for k in range(5)]                     # :745-749
    L1 = (L1_c[0] + L1_c[1] + L1_c[2] + L1_c[3] + L1_G1) / 5.0 + L1_c[4] * 10.0         # :751

    ssim_v = [ssim(rescale_01(cyc_yuv[k], per_image), rescale_01(ds[k], per_image), 5.0) for k in range(5)]
    zero = torch.zeros((), dtype=origs[0].dtype)
    sl = [zero if bits[k] else (-torch.log((1.0 + ssim_v[k]) / 2.0)).mean() for k in range(5)]  # :774-778
    ssim_cyc = (sl[0] + sl[1] + sl[2] + sl[3] + sl[4] * 10.0) / 5.0                     # :779

    sp = [(((cyc_yuv[k] * mask) - (ds[k] * mask)) ** 2).mean() for k in range(5)]       # :792-796
    Spec = (sp[0] + sp[1] + sp[2] + sp[3]) / 5.0 + sp[4] * 5.0                          # :806

    content = ((cyc_yuv[4] - ds[0]) ** 2).mean()                                        # :814
    factor = 1.0 / float(2 * 9 * S * S) ** 2                                            # :817
    style = factor * ((gram_matrix(cyc_yuv[4]) - gram_matrix(ds[4])) ** 2).mean()       # :819-821
    NST = 100.0 * style + content                                                       # :824-826

    total_G = (D1_rf + D3_rf) / 6.0 + L1 * 10.0 + ssim_cyc * 10.0 + NST * 10.0          # :829-832
    total_D = (D1_cls + D3_cls) / 6.0 + (D2_rf + D4_rf) / 6.0 + D4_cls * 0.5 + NST * 10.0   # :837-840
    total_C = (D4_cls + NST) * 10.0                                                     # :844

    out.update(dict(
        gen_input=gen_input, gen_Y=gen_Y, gen_rgb=gen_rgb, cyc_Y=cyc_Y, cyc_rgb=cyc_rgb,
        rf_gen=rf_gen, cls_gen=cls_gen, rf_tgt=rf_tgt, cls_tgt=cls_tgt, d3=d3, d4=d4,
        total_Generator_loss=total_G, total_Discriminator_loss=total_D, total_Classification_loss=total_C,
        G_gan_loss=G_gan, G_clsf_loss=G_clsf, L1_loss_Gen=L1, ssim_cyc_loss=ssim_cyc, Spec_loss=Spec,
        content_loss=content, style_loss=style, total_NST_loss=NST, D4_RealFake_cyc=D4_rf,
        D4_classification_loss=D4_cls, D3_RealFake_cyc=D3_rf, D1_RealFake_loss=D1_rf,
        D3_classification_loss=D3_cls, D1_classification_loss=D1_cls, D2_RealFake_target=D2_rf,
        ssim=ssim_v))
    return out


def _trainable(p):
    return [k for k in p if not (k.endswith("in_gamma") or k.endswith("in_beta"))]


def train_step_grads(Gp, Dp, origs, mask, bits, T, d_noise=None, d_keep=None, live_mask=True,
                     per_image=True, clip: bool = True, q=None):
    """:859-871: grads of (total_D + total_Cls) w.r.t. D vars and of total_G w.r.t. G vars, then
    clip_by_value(+-1).  Returns (losses, gradsG: OrderedDict, gradsD: OrderedDict)."""
    Gp = OrderedDict((k, v.detach().clone().requires_grad_(k in _trainable(Gp))) for k, v in Gp.items())
    Dp = OrderedDict((k, v.detach().clone().requires_grad_(k in _trainable(Dp))) for k, v in Dp.items())
    L = train_step_losses(Gp, Dp, origs, mask, bits, T, d_noise, d_keep, live_mask, per_image, q)
    dnames, gnames = _trainable(Dp), _trainable(Gp)
    gD = torch.autograd.grad(L["total_Discriminator_loss"] + L["total_Classification_loss"],
                             [Dp[k] for k in dnames], retain_graph=True, allow_unused=True)
    gG = torch.autograd.grad(L["total_Generator_loss"], [Gp[k] for k in gnames], allow_unused=True)
    def fin(g, ref):
        g = torch.zeros_like(ref) if g is None else g
        return g.clamp(-1.0, 1.0) if clip else g
    gradsD = OrderedDict((k, fin(g, Dp[k])) for k, g in zip(dnames, gD))
    gradsG = OrderedDict((k, fin(g, Gp[k])) for k, g in zip(gnames, gG))
    L = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in L.items()}
    return L, gradsG, gradsD


def keras_adam_lr(step: int, lr0: float = 2e-5, decay_steps: float = 10000.0, decay_rate: float = 0.95):
    """ExponentialDecay(staircase=False) (:169-171); step = optimizer.iterations BEFORE the update."""
    return lr0 * decay_rate ** (step / decay_steps)


def keras_adam_update(params, grads, m, v, step: int, lr0=2e-5, beta1=0.5, beta2=0.99, eps=1e-7):
    """Keras adam_v2.Adam._resource_apply_dense (:173-174).  step counts from 0; t = step + 1."""
    t = step + 1
    lr = keras_adam_lr(step, lr0)
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in grads:
        g = grads[k]
        m[k] = beta1 * m[k] + (1 - beta1) * g
        v[k] = beta2 * v[k] + (1 - beta2) * g * g
        params[k] = params[k] - lr_t * m[k] / (torch.sqrt(v[k]) + eps)
    return params, m, v


def inference_step(Gp, Sp, rgb, live_mask=True, per_image=True, cyclic: bool = False, stddev_history=None):
    """test.py:218-250: standardise -> SpecSeg mask -> G1 (slot 0 = Y, ED one-hot) -> yuv->rgb.
    cyclic=True adds test.py:252-284: five more generator passes whose non-target slots all carry `gen_rgb[..., 0]` (the R channel,
    named orig_Ych in the reference), slot k zeroed and one-hot plane k set; outputs re-joined with the image's CbCr.
    gen_rgb_output (test.py:249, ShmGANwithSSpecSeg.py:551) = yuv_to_rgb(gen_YCbCr * mean(stddev_arr) * 255), stddev_arr = every scale appended so far
    (`stddev_history` = the earlier ones)."""
    yuv, scale = per_image_standardization(rgb_to_yuv(rgb), per_image)
    Y = yuv[..., 0:1]
    mask = specseg_forward(Sp, Y)
    gmask = mask if live_mask else None
    z = torch.zeros_like(Y)
    o = torch.ones_like(Y)
    gen_input = torch.cat([Y, z, z, z, z, z, z, z, z, o], dim=3)
    gen_Y = generator_forward(Gp, gen_input, gmask)
    cbcr = yuv[..., 1:]
    gen_ycc = torch.cat([gen_Y, cbcr], dim=3)
    gen_rgb = yuv_to_rgb(gen_ycc)
    hist = [scale.reshape(-1)] + ([] if stddev_history is None else [h.reshape(-1).to(scale.dtype) for h in stddev_history])
    avg_std = torch.cat(hist).mean()
    out = dict(mask=mask, gen_Y=gen_Y, gen_rgb=gen_rgb, yuv=yuv, scale=scale,
               gen_rgb_output=yuv_to_rgb(gen_ycc * avg_std * 255.0))
    if cyclic:
        R = gen_rgb[..., 0:1]                                                           # test.py:252
        cyc = []
        for k in range(5):
            planes = [z if j == k else R for j in range(5)]                             # test.py:260-264
            onehot = [o if j == k else z for j in range(5)]                             # test.py:271-275
            cy = generator_forward(Gp, torch.cat(planes + onehot, dim=3), gmask)        # test.py:280-284
            cyc.append(yuv_to_rgb(torch.cat([cy, cbcr], dim=3)))                        # test.py:286-297
        out["cyc_rgb"] = cyc
    return out
