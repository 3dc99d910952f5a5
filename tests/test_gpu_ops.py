"""GPU parity: bandwidth / reduction kernels (instance norm, pooling, preprocessing, losses, Adam) vs the CPU oracle."""
import math

import pytest
import torch

import oracle as O
from _util import F64, bf16_round, dev, rand, randn, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


# ---- instance norm ---------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 8, 12, 8), (3, 4, 4, 512), (2, 32, 32, 16)])
def test_inorm_fwd_pool_add(shape):
    from shmgan_b200 import ops
    N, H, W, C = shape
    x = randn(shape, 1) * 2 + 0.5
    gamma, beta = 1 + 0.1 * randn((C,), 2), 0.02 * randn((C,), 3)
    add = randn((1, H, W, C), 4)
    want = O.instance_norm(x, gamma, beta)
    xd = dev(x)
    sums = ops.inorm_stats(xd)
    ref_s = torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=-1)
    assert rel_err(sums, ref_s) < 1e-6
    y, _ = ops.inorm_apply(xd, sums, dev(gamma), dev(beta))
    assert rel_err(y, want) < TOL
    # fused: + broadcast add, written into the upper half of a concat buffer, + AvgPool2 of the un-added value
    cat = torch.zeros((N, H, W, 2 * C), device="cuda")
    _, pooled = ops.inorm_apply(xd, sums, dev(gamma), dev(beta), add=dev(add), out=cat[..., C:], pooled=True)
    assert rel_err(cat[..., C:], want + add) < TOL
    assert rel_err(pooled, O.avg_pool2(want)) < TOL
    assert float(cat[..., :C].abs().max()) == 0.0


@pytest.mark.parametrize("act", [1, 0])
def test_inorm_bwd(act):
    from shmgan_b200 import ops
    N, H, W, C = 2, 8, 8, 16
    pre = randn((N, H, W, C), 5).requires_grad_()
    gamma, beta = 1 + 0.1 * randn((C,), 6), 0.02 * randn((C,), 7)
    z = O.leaky_relu(pre) if act == 1 else pre
    y = O.instance_norm(z, gamma, beta)
    dyA, dyP = randn((N, H, W, C), 8), randn((N, H // 2, W // 2, C), 9)
    loss = (y * dyA).sum() + (O.avg_pool2(y) * dyP).sum()
    want, = torch.autograd.grad(loss, pre)
    zd = dev(z)
    sums = ops.inorm_stats(zd)
    got = ops.inorm_bwd(zd, sums, dev(gamma), dev(dyA), dev(dyP), act=act)
    assert rel_err(got, want) < TOL
    want_a, = torch.autograd.grad((O.instance_norm(O.leaky_relu(pre) if act == 1 else pre, gamma, beta) * dyA).sum(), pre)
    assert rel_err(ops.inorm_bwd(zd, sums, dev(gamma), dev(dyA), None, act=act), want_a) < TOL



# ---- bf16 fast paths (8 channels per thread, 4 pixels in flight) --------------------------------------
BF = torch.bfloat16


@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (3, 8, 24, 128), (1, 32, 32, 8), (2, 4, 4, 1024), (2, 18, 10, 256)])
def test_inorm_bf16_fast_fwd(shape):
    """inorm_stats8 / inorm_apply8 (+pool, +broadcast add, concat-slice output) on bf16 tensors vs the oracle on the same
    bf16-rounded inputs; the only error left is the bf16 rounding of the outputs (<= 2^-8 relative)."""
    from shmgan_b200 import ops
    N, H, W, C = shape
    x = bf16_round(randn(shape, 1) * 2 + 0.5)
    gamma, beta = 1 + 0.1 * randn((C,), 2), 0.02 * randn((C,), 3)
    add = bf16_round(randn((1, H, W, C), 4))
    want = O.instance_norm(x, gamma, beta)
    # x lives in a channel slice of a wider buffer (pixel stride 2C)
    xbuf = torch.full((N, H, W, 2 * C), 3.0, device="cuda", dtype=BF)
    xbuf[..., :C] = dev(x, BF)
    xd = xbuf[..., :C]
    sums = ops.inorm_stats(xd)
    ref_s = torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=-1)
    assert rel_err(sums, ref_s) < 1e-5
    y, _ = ops.inorm_apply(xd, sums, dev(gamma), dev(beta))
    assert rel_err(y, want) < 8e-3
    cat = torch.zeros((N, H, W, 2 * C), device="cuda", dtype=BF)
    _, pooled = ops.inorm_apply(xd, sums, dev(gamma), dev(beta), add=dev(add, BF), out=cat[..., C:], pooled=True)
    assert rel_err(cat[..., C:], want + add) < 8e-3
    assert rel_err(pooled, O.avg_pool2(want)) < 8e-3
    assert float(cat[..., :C].float().abs().max()) == 0.0


@pytest.mark.parametrize("shape,act", [((2, 16, 16, 64), 1), ((2, 8, 8, 128), 0), ((1, 24, 8, 512), 1), ((3, 6, 10, 8), 2)])
def test_inorm_bf16_fast_bwd_with_bias_grad(shape, act):
    from shmgan_b200 import ops
    N, H, W, C = shape
    pre = randn(shape, 5)
    zv = O.leaky_relu(pre) if act == 1 else (torch.relu(pre) if act == 2 else pre)
    z = bf16_round(zv).requires_grad_()
    gamma = 1 + 0.1 * randn((C,), 6)
    beta = 0.02 * randn((C,), 7)
    y = O.instance_norm(z, gamma, beta)
    dyA, dyP = bf16_round(randn(shape, 8)), bf16_round(randn((N, H // 2, W // 2, C), 9))
    loss = (y * dyA).sum() + (O.avg_pool2(y) * dyP).sum()
    dz, = torch.autograd.grad(loss, z)
    slope = {1: torch.where(z > 0, 1.0, 0.2), 2: (z > 0).to(F64), 0: torch.ones_like(z)}[act]
    want = (dz * slope).detach()
    zd = dev(z.detach(), BF)
    sums = ops.inorm_stats(zd)
    db = torch.full((C,), 0.5, device="cuda")
    dxbuf = torch.full((N, H, W, 2 * C), -1.0, device="cuda", dtype=BF)
    got = ops.inorm_bwd(zd, sums, dev(gamma), dev(dyA, BF), dev(dyP, BF), act=act, dx=dxbuf[..., C:], dbias=db)
    assert rel_err(got, want) < 1e-2
    # the column sums cancel (exactly, without an activation): compare on the scale of the summed magnitudes
    scale = float(want.abs().sum(dim=(0, 1, 2)).max())
    assert float(((db - 0.5).double().cpu() - want.sum(dim=(0, 1, 2))).abs().max()) < 2e-3 * scale
    assert float((dxbuf[..., :C].float() + 1.0).abs().max()) == 0.0
    # gradient arriving through the direct branch only (the per-pixel kernels; the call above used the 2x2-quad kernels)
    dz_a, = torch.autograd.grad((O.instance_norm(z, gamma, beta) * dyA).sum(), z)
    got_a = ops.inorm_bwd(zd, sums, dev(gamma), dev(dyA, BF), None, act=act)
    assert rel_err(got_a, (dz_a * slope).detach()) < 1e-2


def test_act_bwd_bf16_fast_with_bias_grad():
    from shmgan_b200 import ops
    N, H, W, C = 3, 10, 6, 64
    y, dy = bf16_round(randn((N, H, W, C), 16)), bf16_round(randn((N, H, W, C), 17))
    want = dy * torch.where(y > 0, 1.0, 0.2)
    db = torch.zeros((C,), device="cuda")
    ybuf = torch.zeros((N, H, W, 2 * C), device="cuda", dtype=BF)
    ybuf[..., :C] = dev(y, BF)
    got = ops.act_bwd(dev(dy, BF), ybuf[..., :C], 1, dbias=db)
    assert rel_err(got, want) < 8e-3
    assert rel_err(db, want.sum(dim=(0, 1, 2))) < 5e-3


def test_bn_eval_maxpool_actbwd_groupsum():
    from shmgan_b200 import ops
    N, H, W, C = 2, 8, 8, 16
    x = randn((N, H, W, C), 10)
    g, b, m, v = 1 + 0.2 * rand((C,), 11), randn((C,), 12, 0.1), randn((C,), 13, 0.1), 0.5 + rand((C,), 14)
    want = (x - m) * torch.rsqrt(v + O.BN_EPS) * g + b
    cat = torch.zeros((N, H, W, 2 * C), device="cuda")
    _, pooled = ops.bn_eval(dev(x), dev(g), dev(b), dev(m), dev(v), out=cat[..., C:], pooled=True)
    assert rel_err(cat[..., C:], want) < TOL
    assert rel_err(pooled, O.max_pool(want, 2)) < TOL
    mask = rand((N, 32, 32, 1), 15)
    assert torch.equal(ops.maxpool(dev(mask), 16).cpu(), O.max_pool(mask.float(), 16))
    assert torch.equal(ops.maxpool(dev(mask), 2).cpu(), O.max_pool(mask.float(), 2))
    y = randn((N, H, W, C), 16)
    dy = randn((N, H, W, C), 17)
    assert rel_err(ops.act_bwd(dev(dy), dev(y), 1), dy * torch.where(y > 0, 1.0, 0.2)) < 1e-6
    assert rel_err(ops.act_bwd(dev(dy), dev(y), 2), dy * (y > 0)) < 1e-6
    src = randn((6, 4, 4, 8), 18)
    dst = torch.ones((2, 4, 4, 8), device="cuda")
    ops.group_sum(dev(src), 2, dst, accumulate=True)
    assert rel_err(dst, 1 + src.reshape(3, 2, 4, 4, 8).sum(0)) < 1e-6
    a, bb = randn((2, 4, 4, 3), 19), randn((2, 4, 4, 3), 20)
    assert rel_err(ops.add(dev(a), dev(bb)), a + bb) < 1e-6


def test_dense_head():
    from shmgan_b200 import ops
    B, K, J = 3, 4 * 4 * 64, 5
    x, w, dout = randn((B, 4, 4, 64), 21), randn((K, J), 22, 0.02), randn((B, J), 23)
    want = x.reshape(B, -1) @ w
    assert rel_err(ops.dense_fwd(dev(x), dev(w)), want) < TOL
    assert rel_err(ops.dense_dgrad(dev(dout), dev(w), dev(x)), (dout @ w.T).reshape(x.shape)) < TOL
    dw = torch.full((K, J), 0.25, device="cuda")
    ops.dense_wgrad(dev(x), dev(dout), dw)
    assert rel_err(dw - 0.25, x.reshape(B, -1).T @ dout) < TOL


# ---- polarimetric preprocessing ---------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [0, 1, 5, 4 * 37 * 53 * 3, 2 * 256 * 256 * 3])
def test_pseudo_diffuse_min4_bit_exact(dtype, n):
    """calculate_estimate_diffuse (utils.py:102-106): bit-exact in every dtype, ragged and empty sizes included."""
    from shmgan_b200 import ops
    g = torch.Generator().manual_seed(n + 1)
    if dtype == torch.uint8:
        imgs = [torch.randint(0, 256, (n,), generator=g, dtype=torch.uint8) for _ in range(4)]
    else:
        imgs = [torch.rand((n,), generator=g).to(dtype) for _ in range(4)]
    want = O.pseudo_diffuse_min4(*imgs)
    got = ops.pseudo_diffuse_min4(*[t.cuda() for t in imgs])
    assert torch.equal(got.cpu(), want)


def test_pseudo_diffuse_properties_full_size():
    """Size-independent properties at cfg2 size (16 x 256 x 256 x 3): idempotence, permutation invariance, lower bound."""
    from shmgan_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    imgs = [torch.rand((16, 256, 256, 3), generator=g, device="cuda") for _ in range(4)]
    ed = ops.pseudo_diffuse_min4(*imgs)
    assert torch.equal(ops.pseudo_diffuse_min4(ed, ed, ed, ed), ed)
    assert torch.equal(ops.pseudo_diffuse_min4(imgs[3], imgs[1], imgs[0], imgs[2]), ed)
    for t in imgs:
        assert bool((ed <= t).all())
    assert torch.equal(ops.pseudo_diffuse_min4(ed, imgs[0], imgs[1], imgs[2]), ed)


def test_yuv_standardize_avgcbcr_assemble_yuv2rgb():
    from shmgan_b200 import ops
    N, S = 3, 16
    rgbs = [rand((N, S, S, 3), 30 + i) for i in range(5)]
    ds = [O.per_image_standardization(O.rgb_to_yuv(r), True) for r in rgbs]
    dd = []
    for r, (want, scale) in zip(rgbs, ds):
        yuv, sc = ops.yuv_standardize(dev(r))
        assert rel_err(yuv, want) < TOL
        assert rel_err(sc, scale.reshape(N)) < TOL
        dd.append(yuv)
    # constant image: std = 0 -> divisor clamps at 1/256 (ShmGANwithSSpecSeg.py:1299)
    const = torch.zeros((1, S, S, 3), dtype=F64)
    yuv, sc = ops.yuv_standardize(dev(const))
    assert float(sc[0]) == pytest.approx(1.0 / 256.0) and float(yuv.abs().max()) == 0.0
    avg = ops.avg_cbcr(dd)
    want_avg = sum(d[0][..., 1:] for d in ds) / 5.0
    assert rel_err(avg, want_avg) < TOL
    # G1 input assembly (:509-531) with drop bits (1,0,1,0,0)
    bits = [True, False, True, False, False]
    Y = [d[0][..., 0:1] for d in ds]
    want_in = O.assemble_g1_input(Y, bits)
    out = torch.empty((N, S, S, 10), device="cuda")
    ops.assemble_input([None if bits[k] else dd[k] for k in range(5)], [3] * 5, 4, out)
    assert rel_err(out, want_in) < TOL
    # cyclic inputs (:576-594)
    genY = randn((N, S, S, 1), 40)
    gd = dev(genY)
    for k, want_c in enumerate(O.assemble_cyclic_inputs(Y, genY, bits)):
        srcs, lds = [], []
        for j in range(5):
            if j == k:
                srcs.append(None); lds.append(0)
            elif bits[j]:
                srcs.append(gd); lds.append(1)
            else:
                srcs.append(dd[j]); lds.append(3)
        outb = torch.empty((N, S, S, 10), device="cuda", dtype=torch.bfloat16)
        ops.assemble_input(srcs, lds, k, outb)
        assert rel_err(outb, want_c) < 1e-2
        ops.assemble_input(srcs, lds, k, out)
        assert rel_err(out, want_c) < TOL
    # yuv -> rgb with a batch-broadcast CbCr, and its backward
    cb = want_avg
    Yb = randn((2 * N, S, S, 1), 41)
    want_rgb = O.yuv_to_rgb(torch.cat([Yb, cb.repeat(2, 1, 1, 1)], dim=3))
    rgb = torch.empty((2 * N, S, S, 3), device="cuda")
    lp = torch.empty((2 * N, S, S, 3), device="cuda", dtype=torch.bfloat16)
    ops.yuv2rgb(dev(Yb), dev(cb), rgb, lp)
    assert rel_err(rgb, want_rgb) < TOL and rel_err(lp, want_rgb) < 1e-2
    drgb = randn((2 * N, S, S, 3), 42)
    dY = torch.ones((2 * N, S, S, 1), device="cuda")
    ops.yuv2rgb_bwd(dev(drgb), None, dY, accumulate=True)
    assert rel_err(dY, 1 + drgb.sum(dim=3, keepdim=True)) < TOL
    dgen = torch.zeros((N, S, S, 1), device="cuda")
    din = randn((N, S, S, 10), 43)
    ops.assemble_bwd(dev(din), [0, 2], dgen)
    assert rel_err(dgen, din[..., 0:1] + din[..., 2:3]) < TOL


# ---- losses -----------------------------------------------------------------------------------------
def _slot():
    return torch.zeros(1, device="cuda")


def test_elementwise_losses_and_ce():
    from shmgan_b200 import losses as LS, ops
    import ctypes as C
    a = randn((3, 8, 8, 1), 50).requires_grad_()
    want = ((a - 0.9) ** 2).mean()
    g, = torch.autograd.grad(want * 0.7, a)
    s, ad = _slot(), dev(a)
    da = torch.ones_like(ad)
    LS.lsgan(ad, 0.9, C.c_void_p(s.data_ptr()), 2.0, da, 0.7, accumulate=True)
    assert float(s) == pytest.approx(2.0 * float(want), rel=1e-5)
    assert rel_err(da, 1 + g) < TOL
    b = randn((3, 8, 8, 1), 51)
    want = (a - b).abs().mean()
    g, = torch.autograd.grad(want, a)
    s = _slot()
    LS.l1(ad, dev(b), C.c_void_p(s.data_ptr()), 1.0, da, 1.0, accumulate=False)
    assert float(s) == pytest.approx(float(want), rel=1e-5) and rel_err(da, g) < TOL
    logits = randn((4, 5), 52).requires_grad_()
    lab = torch.tensor([[0, 0, 0, 0, 0.9]], dtype=F64)
    want = O.softmax_ce(lab, logits).mean()
    g, = torch.autograd.grad(want * 3.0, logits)
    s = _slot()
    dl = torch.zeros((4, 5), device="cuda")
    LS.softmax_ce(dev(logits), lab[0].tolist(), C.c_void_p(s.data_ptr()), 1.0, dl, 3.0)
    assert float(s) == pytest.approx(float(want), rel=1e-5) and rel_err(dl, g) < TOL


def test_content_style_ssim_spec_losses():
    from shmgan_b200 import losses as LS
    import ctypes as C
    N, S = 2, 32
    Y = (randn((N, S, S, 1), 60) * 0.5 + 1.0).requires_grad_()
    cbcr = randn((N, S, S, 2), 61) * 0.3
    ref = randn((N, S, S, 3), 62) * 0.5 + 0.5
    mask = rand((N, S, S, 1), 63)
    img = torch.cat([Y, cbcr], dim=3)
    Yd, cd, rd = dev(Y), dev(cbcr), dev(ref)

    want = ((img - ref) ** 2).mean()
    g, = torch.autograd.grad(want * 10.0, Y, retain_graph=True)
    s, dY = _slot(), torch.zeros((N, S, S, 1), device="cuda")
    LS.mse_ycc(Yd, cd, rd, C.c_void_p(s.data_ptr()), 1.0, dY, 10.0)
    assert float(s) == pytest.approx(float(want), rel=1e-5) and rel_err(dY, g) < TOL

    factor = 1.0 / float(2 * 9 * S * S) ** 2
    want = factor * ((O.gram_matrix(img) - O.gram_matrix(ref)) ** 2).mean()
    g, = torch.autograd.grad(want * 1000.0, Y, retain_graph=True)
    s, dY = _slot(), torch.zeros((N, S, S, 1), device="cuda")
    LS.style(Yd, cd, rd, S, C.c_void_p(s.data_ptr()), 1.0, dY, 1000.0)
    assert float(s) == pytest.approx(float(want), rel=1e-4) and rel_err(dY, g) < 1e-3

    ss = O.ssim(O.rescale_01(img, True), O.rescale_01(ref, True), 5.0)
    want = (-torch.log((1.0 + ss) / 2.0)).mean()
    g, = torch.autograd.grad(want * 2.0, Y, retain_graph=True)
    s, dY = _slot(), torch.zeros((N, S, S, 1), device="cuda")
    got_ss = LS.ssim_term(Yd, cd, rd, C.c_void_p(s.data_ptr()), 1.0, dY, 2.0)
    assert rel_err(got_ss, ss) < 1e-4
    assert float(s) == pytest.approx(float(want), rel=1e-4)
    assert rel_err(dY, g) < 2e-3

    want = (((img * mask) - (ref * mask)) ** 2).mean()
    s = _slot()
    LS.spec(Yd, cd, rd, dev(mask), C.c_void_p(s.data_ptr()), 1.0)
    assert float(s) == pytest.approx(float(want), rel=1e-5)


def test_clip_adam_matches_keras_update():
    from shmgan_b200 import _lib as L
    import ctypes as C
    n = 1003
    p = {"w": randn((n,), 70)}
    g = {"w": randn((n,), 71) * 2.0}            # some |g| > 1 so the clip is exercised
    m = {"w": torch.zeros(n, dtype=F64)}
    v = {"w": torch.zeros(n, dtype=F64)}
    pd, md, vd = dev(p["w"]), dev(m["w"]), dev(v["w"])
    for step in range(3):
        gc = {"w": g["w"].clamp(-1, 1)}
        p, m, v = O.keras_adam_update(p, gc, m, v, step, 2e-5, 0.5, 0.99, 1e-7)
        t = step + 1
        lr_t = O.keras_adam_lr(step) * math.sqrt(1 - 0.99 ** t) / (1 - 0.5 ** t)
        L.call("shm_clip_adam", C.c_void_p(pd.data_ptr()), C.c_void_p(dev(g["w"]).data_ptr()), C.c_void_p(md.data_ptr()),
               C.c_void_p(vd.data_ptr()), n, lr_t, 0.5, 0.99, 1e-7, 1.0, 1.0, torch.cuda.current_stream().cuda_stream)
    assert rel_err(pd, p["w"]) < 1e-6 and rel_err(md, m["w"]) < 1e-5 and rel_err(vd, v["w"]) < 1e-5


def test_rng_moments():
    from shmgan_b200 import ops
    x = ops.rng_normal((1 << 20,), 7, 0, 0.1, torch.float32)
    assert abs(float(x.mean())) < 1e-3 and float(x.std()) == pytest.approx(0.1, rel=1e-2)
    k = ops.rng_keep((1 << 20,), 7, 1 << 20, 0.8, torch.float32)
    assert float(k.mean()) == pytest.approx(0.8, abs=2e-3)
    assert torch.equal(ops.rng_normal((1000,), 7, 0, 0.1, torch.float32), x[:1000])


# ---- tensor-core first-layer padding and the 1x1 -> 1 channel bandwidth kernels -------------------------
def test_pad64_and_padded_forms():
    from shmgan_b200 import ops
    N, S = 2, 16
    for C, seed in ((1, 80), (3, 81), (10, 82)):
        x = randn((N, S, S, C), seed)
        out = ops.pad64(dev(x))
        assert out.shape == (N, S, S, 64) and out.dtype == torch.bfloat16
        assert torch.equal(out[..., :C].float().cpu(), x.float().bfloat16().float())
        assert float(out[..., C:].abs().max()) == 0.0
    # padded generator-input assembly == plain assembly in the first 10 channels, zeros after
    ds = [dev(randn((N, S, S, 3), 90 + i)) for i in range(5)]
    plain = torch.empty((N, S, S, 10), device="cuda", dtype=torch.bfloat16)
    padded = torch.full((N, S, S, 64), 9.0, device="cuda", dtype=torch.bfloat16)
    srcs, lds = [ds[0], None, ds[2], ds[3], None], [3, 0, 3, 3, 0]
    ops.assemble_input(srcs, lds, 1, plain)
    ops.assemble_input(srcs, lds, 1, padded)
    assert torch.equal(padded[..., :10], plain) and float(padded[..., 10:].abs().max()) == 0.0
    # padded yuv->rgb copy and the strided backward
    Y, cb = randn((N, S, S, 1), 95), randn((N, S, S, 2), 96)
    rgb = torch.empty((N, S, S, 3), device="cuda")
    lp = torch.full((N, S, S, 64), 9.0, device="cuda", dtype=torch.bfloat16)
    ops.yuv2rgb(dev(Y), dev(cb), rgb, lp)
    assert torch.equal(lp[..., :3], rgb.bfloat16()) and float(lp[..., 3:].abs().max()) == 0.0
    d = bf16_round(randn((N, S, S, 64), 97))
    dY = torch.zeros((N, S, S, 1), device="cuda")
    ops.yuv2rgb_bwd(None, dev(d, torch.bfloat16), dY, accumulate=False)
    assert rel_err(dY, d[..., :3].sum(dim=3, keepdim=True)) < 1e-6
    din = bf16_round(randn((N, S, S, 64), 98))
    dgen = torch.zeros((N, S, S, 1), device="cuda")
    ops.assemble_bwd(dev(din, torch.bfloat16), [1, 4], dgen)
    assert rel_err(dgen, din[..., 1:2] + din[..., 4:5]) < 1e-6


@pytest.mark.parametrize("C,act", [(64, 1), (16, 3), (256, 1), (8, 0)])
def test_pw1_kernels(C, act):
    from shmgan_b200 import ops
    from _util import oracle_conv
    N, H, W = 2, 13, 11                                   # ragged pixel count: exercises the tail of the last block
    x = bf16_round(randn((N, H, W, C), 100))
    w = randn((1, 1, C, 1), 101, 0.1).float().to(F64)
    b = randn((1,), 102, 0.1).float().to(F64)
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    y = oracle_conv(xr, wr, br, 1, False, act)
    c = ops.Conv("t", 1, 1, C, 1, act=act)
    c.w, c.b = dev(w), dev(b)
    c.dw, c.db = torch.full_like(c.w, 0.5), torch.zeros(1, device="cuda")
    xd = dev(x, torch.bfloat16)
    assert c.pw1_ok(xd)
    yd = c.fwd(xd, tc=True)
    assert rel_err(yd, y) < 1e-2
    if act == 3:
        return
    # backward from the device's own (bf16-rounded) output so that the activation branch is identical
    yq = yd.double().cpu()
    dy = bf16_round(randn(tuple(y.shape), 103))
    g = dy * torch.where(yq > 0, 1.0, 0.2) if act == 1 else dy
    pre = oracle_conv(xr, wr, br, 1, False, 0)
    gx, gw, gb = torch.autograd.grad((pre * g).sum(), [xr, wr, br])
    dx = c.pw1_bwd(xd, dev(dy, torch.bfloat16), yd)
    assert rel_err(dx, gx) < 1e-2
    assert rel_err(c.dw - 0.5, gw) < 1e-3 and rel_err(c.db, gb) < 1e-3


def test_im2col_col2im_k3s2():
    """d1's patch tensor: channel (ky*3+kx)*C + c of output pixel (oy, ox) = x[2oy+ky, 2ox+kx, c] (TF SAME pad (0, 1) at even sizes),
    and col2im as its exact transpose (<im2col(x), g> == <x, col2im(g)>)."""
    from shmgan_b200 import ops
    N, H, W, C = 2, 8, 12, 3
    x = bf16_round(randn((N, H, W, C), 61))
    xp = torch.zeros((N, H + 1, W + 1, C), dtype=F64)
    xp[:, :H, :W] = x
    want = torch.zeros((N, H // 2, W // 2, 64), dtype=F64)
    for ky in range(3):
        for kx in range(3):
            want[..., (ky * 3 + kx) * C:(ky * 3 + kx + 1) * C] = xp[:, ky:ky + H:2, kx:kx + W:2]
    got = ops.im2col_k3s2(dev(x))
    assert got.dtype == torch.bfloat16 and torch.equal(got.double().cpu(), want)
    xbuf = torch.zeros((N, H, W, 8), device="cuda", dtype=torch.bfloat16)       # bf16 source in a wider buffer
    xbuf[..., :C] = dev(x, torch.bfloat16)
    assert torch.equal(ops.im2col_k3s2(xbuf[..., :C]).double().cpu(), want)
    # the strip kernel of the discriminator's dense bf16 RGB input (W/2 a multiple of 32): one and two strips per row, odd image count
    for (n2, h2, w2) in ((1, 4, 64), (3, 6, 128), (2, 2, 192)):
        x2 = bf16_round(randn((n2, h2, w2, 3), 63))
        xp2 = torch.zeros((n2, h2 + 1, w2 + 1, 3), dtype=F64)
        xp2[:, :h2, :w2] = x2
        want2 = torch.zeros((n2, h2 // 2, w2 // 2, 64), dtype=F64)
        for ky in range(3):
            for kx in range(3):
                want2[..., (ky * 3 + kx) * 3:(ky * 3 + kx + 1) * 3] = xp2[:, ky:ky + h2:2, kx:kx + w2:2]
        assert torch.equal(ops.im2col_k3s2(dev(x2, torch.bfloat16)).double().cpu(), want2), (n2, h2, w2)
    # ... and the strip form of col2im (dense bf16 dx, W a multiple of 64) against a scatter-add restatement
    for (n2, h2, w2) in ((1, 4, 64), (3, 6, 128)):
        g2 = bf16_round(randn((n2, h2 // 2, w2 // 2, 64), 64))
        acc = torch.zeros((n2, h2 + 1, w2 + 1, 3), dtype=F64)
        for ky in range(3):
            for kx in range(3):
                acc[:, ky:ky + h2:2, kx:kx + w2:2] += g2[..., (ky * 3 + kx) * 3:(ky * 3 + kx + 1) * 3]
        got2 = ops.col2im_k3s2(dev(g2, torch.bfloat16), 3, h2, w2, torch.bfloat16)
        assert got2.dtype == torch.bfloat16 and rel_err(got2, acc[:, :h2, :w2]) < 1e-2, (n2, h2, w2)
    g = bf16_round(randn((N, H // 2, W // 2, 64), 62))
    dx = ops.col2im_k3s2(dev(g, torch.bfloat16), C, H, W, torch.float32)
    lhs = float((want * g).sum())
    rhs = float((x * dx.double().cpu()).sum())
    assert abs(lhs - rhs) < 1e-3 * (abs(lhs) + 1.0)
