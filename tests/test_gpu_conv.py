"""GPU parity: convolution kernels (exact-fp32 SIMT path and tcgen05 bf16 path) through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): 1e-3 relative in fp32 mode, 2e-2 in bf16 mode; the fp32 kernels are held to 1e-4 here.
"""
import pytest
import torch

import oracle as O
from _util import F64, bf16_round, dev, oracle_conv, randn, rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _mk_conv(cin, cout, k, stride, transposed, act, bias, w, b):
    from shmgan_b200 import ops
    c = ops.Conv("t", k, k, cin, cout, stride=stride, transposed=transposed, act=act, bias=bias)
    c.w = dev(w)
    c.b = dev(b) if bias else None
    c.dw = torch.zeros_like(c.w)
    c.db = torch.zeros(cout, device="cuda") if bias else None
    return c


# (N, H, W, Cin, Cout, k, stride, transposed, act, bias)
SIMT_CASES = [
    (2, 16, 16, 10, 64, 3, 1, False, 1, True),      # enc1a class
    (2, 16, 16, 3, 64, 3, 2, False, 1, False),      # d1 class: TF SAME pad (0, 1)
    (1, 12, 20, 64, 1, 1, 1, False, 1, True),       # generator output 1x1 -> 1 channel
    (2, 8, 8, 32, 16, 3, 2, True, 1, True),         # Conv2DTranspose k3 s2 (generator up path)
    (2, 8, 8, 32, 16, 2, 2, True, 0, True),         # Conv2DTranspose k2 s2 (SpecSeg)
    (1, 9, 7, 5, 7, 3, 1, False, 2, True),          # ragged / odd sizes
    (2, 7, 9, 4, 6, 3, 2, False, 0, True),          # odd input, stride 2: pad (1, 1)
    (1, 8, 8, 256, 1, 3, 1, False, 1, False),       # discriminator real/fake head class
    (2, 16, 16, 64, 128, 3, 1, False, 1, True),
    (1, 4, 4, 1, 32, 3, 1, False, 1, True),         # attention first conv: single input channel
    (1, 16, 16, 16, 1, 1, 1, False, 3, True),       # SpecSeg sigmoid head
]


@pytest.mark.parametrize("case", SIMT_CASES)
def test_conv_fp32_fwd_dgrad_wgrad(case):
    N, H, W, Cin, Cout, k, s, tr, act, bias = case
    x = randn((N, H, W, Cin), 1)
    wshape = (k, k, Cout, Cin) if tr else (k, k, Cin, Cout)
    w = randn(wshape, 2, 0.1)
    b = randn((Cout,), 3, 0.1)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    pre = oracle_conv(xr, wr, br if bias else None, s, tr, 0)
    want = oracle_conv(x, w, b if bias else None, s, tr, act)
    dy = randn(tuple(pre.shape), 4)
    grads = torch.autograd.grad((pre * dy).sum(), [xr, wr] + ([br] if bias else []))

    c = _mk_conv(Cin, Cout, k, s, tr, act, bias, w, b)
    xd = dev(x)
    y = c.fwd(xd, tc=False)
    assert rel_err(y, want) < FP32_TOL
    dyd = dev(dy)
    dx = c.dgrad(dyd, xd.shape, tc=False)
    assert rel_err(dx, grads[0]) < FP32_TOL
    # wgrad accumulates on top of what is already there
    c.dw.fill_(0.5)
    c.wgrad(xd, dyd, tc=False)
    assert rel_err(c.dw - 0.5, grads[1]) < FP32_TOL
    if bias:
        assert rel_err(c.db, grads[2]) < FP32_TOL


def test_conv_fp32_channel_slices():
    """x read from and y written into channel slices of wider NHWC buffers (concat never materialised)."""
    N, H, W, Cin, Cout = 2, 8, 8, 16, 32
    x, w, b = randn((N, H, W, Cin), 5), randn((3, 3, Cin, Cout), 6, 0.1), randn((Cout,), 7, 0.1)
    want = oracle_conv(x, w, b, 1, False, 1)
    c = _mk_conv(Cin, Cout, 3, 1, False, 1, True, w, b)
    xbuf = torch.full((N, H, W, 2 * Cin), 7.0, device="cuda")
    xbuf[..., Cin:] = dev(x)
    ybuf = torch.full((N, H, W, 3 * Cout), -3.0, device="cuda")
    c.fwd(xbuf[..., Cin:], ybuf[..., Cout:2 * Cout], tc=False)
    torch.cuda.synchronize()
    assert rel_err(ybuf[..., Cout:2 * Cout], want) < FP32_TOL
    assert float((ybuf[..., :Cout] + 3.0).abs().max()) == 0.0 and float((ybuf[..., 2 * Cout:] + 3.0).abs().max()) == 0.0


def test_conv_transpose_impulse_alignment():
    """Conv2DTranspose(k3, s2, SAME): out[2i+k] += x[i] w[k], cropped at the END."""
    x = torch.zeros(1, 4, 4, 1, dtype=F64)
    x[0, 1, 2, 0] = 1.0
    w = torch.arange(1, 10, dtype=F64).reshape(3, 3, 1, 1)
    c = _mk_conv(1, 1, 3, 2, True, 0, False, w, None)
    y = c.fwd(dev(x), tc=False).cpu()[0, :, :, 0]
    want = torch.zeros(8, 8)
    want[2:5, 4:7] = w[:, :, 0, 0].float()
    assert torch.equal(y, want)


def test_conv_bad_args_fail_loudly():
    from shmgan_b200 import _lib as L
    import ctypes as C
    d = L.ConvDesc(1, 8, 8, 4, 4, 5, 5, 1, 0, 0, 4, 4, 0, 0)      # 5x5 kernel: unsupported
    with pytest.raises(L.ShmError):
        L.call("shm_conv2d_fwd", C.byref(d), None, None, None, None, None)


# ---- tcgen05 path --------------------------------------------------------------------------------
# (N, H, W, Cin, Cout, k, stride, transposed, act, bias)
TC_CASES = [
    (2, 16, 16, 64, 64, 3, 1, False, 1, True),      # box 16x8x1
    (2, 8, 8, 128, 128, 3, 1, False, 1, True),      # box 8x8x2 (two images per tile)
    (8, 4, 4, 64, 128, 3, 1, False, 1, True),       # box 4x4x8
    (1, 16, 16, 512, 512, 1, 1, False, 1, True),    # bottleneck 1x1
    (2, 32, 32, 64, 128, 3, 2, False, 1, False),    # discriminator block, stride 2
    (2, 8, 8, 128, 64, 3, 2, True, 1, True),        # Conv2DTranspose up path
    (1, 4, 128, 64, 64, 3, 1, False, 1, True),      # box 128x1
    (1, 2, 256, 64, 64, 3, 1, False, 0, True),      # two tiles per row
    (2, 16, 16, 128, 64, 3, 1, False, 1, True),     # dec-a class (Cin = 2 * Cout)
    (1, 16, 16, 256, 256, 3, 1, False, 2, True),
    (2, 32, 32, 64, 64, 3, 1, False, 1, True),      # halo kernel (resident weights, one halo tile for all 9 taps), 8 tiles/image
    (1, 48, 16, 128, 64, 3, 1, False, 1, True),     # halo kernel, two k-chunks, non-square
    (3, 16, 16, 64, 128, 3, 1, False, 0, True),     # halo kernel BN = 128 (its dgrad: two k-chunks, BN = 64)
    (2, 32, 16, 128, 128, 3, 1, False, 1, True),    # halo wgrad MODE 1 (128 ci x 128 co x filter row units)
    (3, 16, 32, 256, 128, 3, 1, False, 1, True),    # halo wgrad MODE 1, two ci blocks, split over tiles
    (5, 32, 32, 64, 64, 3, 1, False, 1, True),      # halo wgrad MODE 0, uneven split of 40 tiles
    (2, 32, 32, 128, 128, 3, 1, False, 1, True),    # big halo kernel (32x8 tile = two accumulators per weight tile), fwd and dgrad
    (1, 64, 16, 256, 128, 3, 1, False, 2, True),    # big halo kernel, four k-chunks (A ring wraps), dgrad has two n-tiles
    (2, 32, 8, 64, 256, 3, 1, False, 0, True),      # big halo kernel fwd with two n-tiles, single k-chunk
    (10, 64, 64, 128, 256, 3, 1, False, 1, True),   # big halo kernel, 320 work items: persistent loop, accumulator double buffering
    (3, 32, 8, 128, 128, 3, 1, False, 1, True),     # big halo kernel, CTA-pair form with an ODD tile count (the last pair's second tile is a dummy)
    (150, 32, 8, 128, 128, 3, 1, False, 0, True),   # CTA-pair form, 75 pairs on 74 clusters: one cluster wraps around, operand rings wrap
    (1, 32, 8, 192, 128, 3, 1, False, 1, True),     # single tile: falls back to the single-CTA big kernel
    (2, 16, 16, 128, 64, 3, 2, True, 1, True),      # scatter kernel: Conv2DTranspose fwd, four parity accumulators x 64 columns
    (1, 32, 16, 256, 128, 3, 2, True, 1, True),     # scatter kernel, 4 x 128 columns (single accumulator set), four k-chunks
    (3, 16, 8, 64, 128, 2, 2, True, 0, True),       # scatter kernel, k2 s2 transposed conv (SpecSeg up path)
    (2, 64, 64, 128, 256, 3, 2, False, 1, False),   # stride-2 conv: dgrad through the scatter kernel (K = 256, N = 128)
    (20, 32, 32, 128, 128, 3, 2, True, 1, True),    # scatter kernel, 160 work items on one accumulator set (persistent loop)
    (40, 16, 16, 64, 64, 3, 2, True, 1, True),      # scatter kernel, 80 x ... double-buffered sets across many items
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_fwd_dgrad_wgrad(case):
    from shmgan_b200 import ops
    import ctypes as C
    N, H, W, Cin, Cout, k, s, tr, act, bias = case
    x = bf16_round(randn((N, H, W, Cin), 11))
    wshape = (k, k, Cout, Cin) if tr else (k, k, Cin, Cout)
    w = bf16_round(randn(wshape, 12, 0.05))
    b = randn((Cout,), 13, 0.1).float().to(F64)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    pre = oracle_conv(xr, wr, b if bias else None, s, tr, 0)
    want = oracle_conv(x, w, b if bias else None, s, tr, act)
    dy = bf16_round(randn(tuple(pre.shape), 14))
    gx, gw = torch.autograd.grad((pre * dy).sum(), [xr, wr])

    c = _mk_conv(Cin, Cout, k, s, tr, act, bias, w, b)
    xd = dev(x, torch.bfloat16)
    d = c.desc(N, H, W, Cin, Cout, ops.BF16)
    assert c.tc_ok(d), "case must be servable by the tensor-core path"
    y = c.fwd(xd, tc=True, version=1)
    assert y.dtype == torch.bfloat16
    assert rel_err(y, want) < BF16_TOL
    dyd = dev(dy, torch.bfloat16)
    dx = c.dgrad(dyd, xd.shape, tc=True, version=1)
    assert rel_err(dx, gx) < BF16_TOL
    c.wgrad(xd, dyd, tc=True)
    assert rel_err(c.dw, gw) < BF16_TOL


def test_conv_tc_halo_channel_slices():
    """Halo kernel reading a channel slice of a wider buffer (the concat input of a decoder block) and writing into one."""
    from shmgan_b200 import ops
    N, H, W, Cin, Cout = 2, 32, 16, 64, 64
    x = bf16_round(randn((N, H, W, Cin), 31))
    w = bf16_round(randn((3, 3, Cin, Cout), 32, 0.05))
    b = randn((Cout,), 33, 0.1).float().to(F64)
    want = oracle_conv(x, w, b, 1, False, 1)
    c = _mk_conv(Cin, Cout, 3, 1, False, 1, True, w, b)
    xbuf = torch.full((N, H, W, 2 * Cin), 7.0, device="cuda", dtype=torch.bfloat16)
    xbuf[..., Cin:] = dev(x, torch.bfloat16)
    ybuf = torch.full((N, H, W, 2 * Cout), -3.0, device="cuda", dtype=torch.bfloat16)
    c.fwd(xbuf[..., Cin:], ybuf[..., :Cout], tc=True, version=1)
    torch.cuda.synchronize()
    assert rel_err(ybuf[..., :Cout], want) < BF16_TOL
    assert float((ybuf[..., Cout:].float() + 3.0).abs().max()) == 0.0


def test_conv_tc_halo_wgrad_channel_slices_and_accumulate():
    """Halo wgrad reading x from a channel slice of a concat buffer and dy from a slice, accumulating on top of dw."""
    N, H, W, Cin, Cout = 2, 32, 16, 128, 64
    x = bf16_round(randn((N, H, W, Cin), 41))
    w = bf16_round(randn((3, 3, Cin, Cout), 42, 0.05))
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    pre = oracle_conv(xr, wr, None, 1, False, 0)
    dy = bf16_round(randn(tuple(pre.shape), 43))
    gw, = torch.autograd.grad((pre * dy).sum(), [wr])
    c = _mk_conv(Cin, Cout, 3, 1, False, 1, False, w, None)
    xbuf = torch.full((N, H, W, Cin + 64), 7.0, device="cuda", dtype=torch.bfloat16)
    xbuf[..., 64:] = dev(x, torch.bfloat16)
    dybuf = torch.full((N, H, W, 2 * Cout), -3.0, device="cuda", dtype=torch.bfloat16)
    dybuf[..., :Cout] = dev(dy, torch.bfloat16)
    c.dw.fill_(0.25)
    c.wgrad(xbuf[..., 64:], dybuf[..., :Cout], tc=True)
    c.wgrad(xbuf[..., 64:], dybuf[..., :Cout], tc=True)
    assert rel_err(c.dw - 0.25, 2.0 * gw) < BF16_TOL


@pytest.mark.parametrize("shape", [(2, 8, 8, 1024), (3, 4, 6, 64), (1, 5, 3, 8)])
def test_conv3x3_to_one_channel_bf16(shape):
    """The discriminator's real/fake head (3x3, C -> 1, no bias, LeakyReLU) as warp-per-pixel dot products: fwd, dgrad, wgrad."""
    N, H, W, C = shape
    x = bf16_round(randn((N, H, W, C), 51))
    w = randn((3, 3, C, 1), 52, 0.05).float().to(F64)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    pre = oracle_conv(xr, wr, None, 1, False, 0)
    want = oracle_conv(x, w, None, 1, False, 1)
    dy = bf16_round(randn(tuple(pre.shape), 53))
    gx, gw = torch.autograd.grad((pre * dy).sum(), [xr, wr])
    c = _mk_conv(C, 1, 3, 1, False, 1, False, w, None)
    xd = dev(x, torch.bfloat16)
    assert c.c3to1_ok(xd)
    y = c.fwd(xd, tc=True, version=1)
    assert y.shape == (N, H, W, 1) and rel_err(y, want) < 8e-3
    dyd = dev(dy, torch.bfloat16)
    dx = c.dgrad(dyd, xd.shape, tc=True, version=1)
    assert rel_err(dx, gx) < 8e-3
    c.dw.fill_(0.5)
    c.wgrad(xd, dyd, tc=True)
    assert rel_err(c.dw - 0.5, gw) < 1e-4


def test_conv_tc_matches_simt_bf16_inputs():
    """Same bf16 inputs through both kernel families: the tcgen05 result differs from the fp32-accumulating SIMT kernel only
    by the bf16 output rounding."""
    N, H, W, Cin, Cout = 2, 16, 16, 64, 64
    x = bf16_round(randn((N, H, W, Cin), 21))
    w = bf16_round(randn((3, 3, Cin, Cout), 22, 0.05))
    b = randn((Cout,), 23, 0.1)
    c = _mk_conv(Cin, Cout, 3, 1, False, 1, True, w, b)
    xd = dev(x, torch.bfloat16)
    y_tc = c.fwd(xd, tc=True, version=1).float()
    y_simt = c.fwd(xd, tc=False).float()
    assert rel_err(y_tc, y_simt) < 1e-2


# ---- thin layers: fewer than 64 reduction / output channels through the halo kernel with 32- / 64-byte pixel rows -----------------
# (N, H, W, cin, cout, seg_pad, cout_dev, act)
THIN_CASES = [
    (2, 32, 32, 1, 16, 16, 16, 2),        # SpecSeg c1a: 1 -> 16 (SWIZZLE_32B rows, 16 accumulator columns)
    (3, 16, 24, 16, 16, 16, 16, 2),       # c1b / c9b
    (2, 32, 16, 16, 32, 16, 32, 2),       # c2a
    (2, 16, 16, 32, 32, 32, 32, 2),       # c2b (SWIZZLE_64B rows)
    (2, 32, 32, 32, 64, 32, 64, 2),       # c3a
    (2, 16, 32, 32, 16, 32, 16, 2),       # c9a on the [up16 | skip16] concat
    (2, 32, 32, 64, 32, 64, 32, 2),       # c8a on the [up32 | skip32] concat (128-byte rows, 32 columns)
    (1, 48, 40, 64, 16, 64, 16, 1),       # 64 -> 16
    (2, 32, 32, 10, 64, 16, 64, 1),       # generator enc1a: 10 -> 64 with the input padded to 16
    (2, 16, 16, 1, 128, 16, 128, 1),      # attention first conv 1 -> 128
    (40, 32, 32, 16, 16, 16, 16, 2),      # 320 tiles: persistent loop, ring wrap, accumulator double buffering
    (2, 32, 32, 32, 32, 32, 64, 2),       # c8b: 32 real output channels in 64 columns
]


@pytest.mark.parametrize("case", THIN_CASES)
def test_conv_tc_thin_layers(case):
    from shmgan_b200 import ops
    N, H, W, cin, cout, seg_pad, cout_dev, act = case
    x = bf16_round(randn((N, H, W, cin), 31))
    w = bf16_round(randn((3, 3, cin, cout), 32, 0.1))
    b = randn((cout,), 33, 0.1).float().to(F64)
    want = oracle_conv(x, w, b, 1, False, act)
    c = ops.PaddedConv("thin", 3, 3, cin, cout, cin, 1, act=act, seg_pad=seg_pad, cout_dev=cout_dev)
    c.w, c.b = dev(w), dev(b)
    assert c.servable(N, H, W)
    xd = ops.pad_channels(dev(x, torch.bfloat16), seg_pad) if cin != seg_pad else dev(x, torch.bfloat16)
    y = c.fwd(xd, None, True, 1)
    assert y.shape == (N, H, W, cout_dev) and y.dtype == torch.bfloat16
    assert rel_err(y[..., :cout], want) < BF16_TOL
    if cout_dev > cout:
        assert float(y[..., cout:].float().abs().max()) == 0.0
    # output placed into a channel slice of a wider buffer (concat placement)
    wide = torch.full((N, H, W, cout_dev + 16), 7.0, dtype=torch.bfloat16, device="cuda")
    c.fwd(xd, wide[..., 16:], True, 1)
    assert torch.equal(wide[..., 16:], y) and float((wide[..., :16].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout,k", [(64, 32, 2), (32, 16, 2), (64, 16, 3)])
def test_conv_transpose_partial_column_store(cin, cout, k):
    """Conv2DTranspose whose real output channels (32 / 16) sit in 64 accumulator columns: only the real ones are stored, into the
    [up | skip] concat buffer, and the skip half is untouched (SpecSeg.py:68,78)."""
    from shmgan_b200 import ops
    N, H, W = 3, 16, 24
    x = bf16_round(randn((N, H, W, cin), 41))
    w = bf16_round(randn((k, k, cout, cin), 42, 0.1))
    b = randn((cout,), 43, 0.1).float().to(F64)
    want = oracle_conv(x, w, b, 2, True, 0)
    c = ops.PaddedConv("up", k, k, cin, cout, cin, 1, stride=2, transposed=True, act=0, nstore=cout)
    c.w, c.b = dev(w), dev(b)
    xd = ops.pad64(dev(x, torch.bfloat16)) if cin < 64 else dev(x, torch.bfloat16)
    cat = torch.full((N, 2 * H, 2 * W, 2 * cout), 3.0, dtype=torch.bfloat16, device="cuda")
    c.fwd(xd, cat[..., :cout], True, 1)
    assert rel_err(cat[..., :cout], want) < BF16_TOL
    assert float((cat[..., cout:].float() - 3.0).abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 32, 32), (3, 64, 16), (1, 16, 8)])
def test_conv_tc_thin_dgrad_of_padded_first_layer(shape):
    """The zero-padded first layer (10 -> 64): its dgrad through the thin halo kernel (K = 64 output channels, N = 16 input columns)
    against the oracle; channels 10..15 of dx belong to weights that do not exist and must be exactly zero."""
    from shmgan_b200 import ops
    N, H, W = shape
    cin, cout = 10, 64
    x = bf16_round(randn((N, H, W, cin), 61))
    w = bf16_round(randn((3, 3, cin, cout), 62, 0.1))
    xr = x.clone().requires_grad_()
    pre = oracle_conv(xr, w, None, 1, False, 0)
    dy = bf16_round(randn(tuple(pre.shape), 63))
    gx, = torch.autograd.grad((pre * dy).sum(), [xr])
    c = _mk_conv(cin, cout, 3, 1, False, 1, False, w, None).enable_pad()
    dx = c.dgrad(dev(dy, torch.bfloat16), (N, H, W, 64), None, True, 1)
    assert dx.shape == (N, H, W, 16)
    assert rel_err(dx[..., :cin], gx) < BF16_TOL
    assert float(dx[..., cin:].float().abs().max()) == 0.0


def test_conv_tc_wgrad_of_first_layer_from_16_channel_input():
    """Weight gradient of the zero-padded first layer (10 -> 64) read from a DENSE 16-channel input: the 64-channel TMA box of the halo
    wgrad kernel runs past the channel extent and is zero-filled, so the result equals the one from the 64-channel padded tensor."""
    from shmgan_b200 import ops
    N, H, W, cin, cout = 3, 32, 32, 10, 64
    x = bf16_round(randn((N, H, W, cin), 71))
    w = bf16_round(randn((3, 3, cin, cout), 72, 0.1))
    wr = w.clone().requires_grad_()
    pre = oracle_conv(x, wr, None, 1, False, 0)
    dy = bf16_round(randn(tuple(pre.shape), 73))
    gw, = torch.autograd.grad((pre * dy).sum(), [wr])
    got = []
    for cpad in (16, 64):
        c = _mk_conv(cin, cout, 3, 1, False, 1, False, w, None).enable_pad()
        c.wgrad(ops.pad_channels(dev(x, torch.bfloat16), cpad), dev(dy, torch.bfloat16), tc=True)
        c.fold_pad_grad()
        assert rel_err(c.dw, gw) < BF16_TOL
        assert float(c.dw_pad[:, :, cin:].abs().max()) == 0.0
        got.append(c.dw.clone())
    assert float((got[0] - got[1]).abs().max()) <= 1e-5 * float(got[1].abs().max())       # fp32 atomics order only


# ---- instance-norm statistics from the convolution epilogue (shm_conv2d_tc_fwd_stats) ----------------------------------------
# (N, H, W, Cin, Cout, k, stride, act, bias, fused?, kernel that serves it)
STATS_CASES = [
    (3, 32, 32, 64, 64, 3, 1, 1, True, True, "halo, TMA-store epilogue (64 -> 64)"),
    (2, 48, 16, 128, 64, 3, 1, 1, True, True, "halo, two k-chunks (dec4a class), direct-store epilogue"),
    (200, 32, 32, 64, 64, 3, 1, 1, True, True, "halo, 1600 tiles: several images and flushes per CTA"),
    (7, 16, 8, 64, 64, 3, 1, 0, False, True, "halo, 7 tiles: fewer tiles than CTAs, no bias, no activation"),
    (3, 16, 16, 64, 128, 3, 1, 1, True, False, "halo BN = 128 (enc2a class): separate pass"),
    (5, 32, 32, 128, 128, 3, 1, 1, True, False, "big halo kernel: separate pass"),
    (4, 64, 64, 64, 128, 3, 2, 1, False, False, "generic kernel, stride 2 (discriminator block): separate pass"),
]


@pytest.mark.parametrize("case", STATS_CASES, ids=[c[-1] for c in STATS_CASES])
def test_conv_tc_fwd_stats_matches_separate_pass(case):
    """Conv.fwd(want_stats=True) must return what shm_inorm_stats computes from the stored tensor (sums of the bf16-rounded outputs) whether the
    statistics come from the convolution's epilogue (64-column halo variants) or from the separate pass, and the stored tensor itself must be
    bit-identical to the plain forward (the tile order of a statistics launch differs, the arithmetic does not)."""
    from shmgan_b200 import ops
    import ctypes as C
    N, H, W, Cin, Cout, k, s, act, bias, fused, _ = case
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn((N, H, W, Cin), device="cuda", generator=g) * 0.7 + 0.1).bfloat16()
    c = ops.Conv("t", k, k, Cin, Cout, stride=s, act=act, bias=bias)
    c.w = (torch.randn((k, k, Cin, Cout), device="cuda", generator=g) * 0.05).bfloat16().float()
    c.b = torch.randn(Cout, device="cuda", generator=g) * 0.3 if bias else None
    d = c.desc(N, H, W, Cin, Cout, ops.BF16)
    assert c.tc_ok(d) and ops.call("shm_conv2d_tc_stats_supported", C.byref(d)) == int(fused)
    y0 = c.fwd(x, tc=True, version=1)
    want = ops.inorm_stats(y0)
    y1, got = c.fwd(x, tc=True, version=1, want_stats=True)
    assert torch.equal(y0, y1)
    yf = y0.double()
    ref = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1)          # [N, C, 2] from the stored values, fp64
    scale = ref.abs().amax(dim=(0, 1), keepdim=True)
    assert float(((got - ref).abs() / scale).max()) < 1e-5                                # fp32 running sums per thread, fp64 across threads
    assert float(((got - want).abs() / scale).max()) < 1e-5
    # a channel-slice destination (the decoder's concat buffer) changes nothing
    wide = torch.zeros((N, y0.shape[1], y0.shape[2], Cout + 64), dtype=torch.bfloat16, device="cuda")
    _, got2 = c.fwd(x, wide[..., 64:], tc=True, version=1, want_stats=True)
    assert torch.equal(wide[..., 64:], y0) and float(((got2 - ref).abs() / scale).max()) < 1e-5
    if not fused:
        with pytest.raises(ops.L.ShmError):
            ops.call("shm_conv2d_tc_fwd_stats", C.byref(d), ops._p(x), ops._p(c.w_tc), ops._p(c.b), ops._p(y1), ops._p(got), ops._stream())


