"""Committed fixtures (tests/golden/*.npz, written by tools/make_golden.py from the CPU oracle in float64).

CPU half: the oracle still reproduces every stored number (a regression pin of the restated TF semantics -- the reference
itself cannot run here, see oracle/__init__.py: PARITY UNPINNED).  GPU half: the C-ABI kernels against the same stored
vectors, without calling the oracle at test time."""
import os

import numpy as np
import pytest
import torch

import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F = torch.float64


def _load(name):
    z = np.load(os.path.join(G, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


CONV_TAGS = ["k3s1", "k3s2", "k3s2_odd", "t_k3s2", "t_k2s2", "k1s1"]


# ---- CPU: oracle vs fixtures ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", CONV_TAGS)
def test_oracle_conv_layers_match_golden(tag):
    z = _load("conv_layers")
    N, H, W, Ci, Co, k, s, tr = [int(v) for v in z[tag + "_cfg"]]
    x, w, b = z[tag + "_x"].requires_grad_(), z[tag + "_w"].requires_grad_(), z[tag + "_b"]
    y = O.conv2d_transpose_same(x, w, b, s) if tr else O.conv2d_same(x, w, b, s)
    assert tuple(y.shape) == tuple(z[tag + "_y"].shape)
    assert _rel(y, z[tag + "_y"]) < 1e-12
    gx, gw = torch.autograd.grad((y * z[tag + "_dy"]).sum(), [x, w])
    assert _rel(gx, z[tag + "_dx"]) < 1e-12 and _rel(gw, z[tag + "_dw"]) < 1e-12


def test_oracle_networks_match_golden():
    z = _load("nets")
    fs, S = int(z["fs"]), int(z["S"])
    Gp = O.init_params(O.generator_param_specs(fs, True), 1, F, randomize_all=True)
    Dp = O.init_params(O.discriminator_param_specs(S, fs, True), 2, F, randomize_all=True)
    assert _rel(O.generator_forward(Gp, z["x"], z["mask"]), z["g_out"]) < 1e-10
    assert _rel(O.generator_forward(Gp, z["x"], None), z["g_out_nomask"]) < 1e-10
    rf, cls = O.discriminator_forward(Dp, z["img"], z["mask"])
    assert _rel(rf, z["d_rf"]) < 1e-10 and _rel(cls, z["d_cls"]) < 1e-10


def test_oracle_prep_match_golden():
    z = _load("prep")
    yuv, _ = O.per_image_standardization(O.rgb_to_yuv(z["img"]), True)
    assert _rel(yuv, z["yuv"]) < 1e-12
    assert _rel(O.yuv_to_rgb(O.rgb_to_yuv(z["img"])), z["rgb_back"]) < 1e-12
    assert torch.equal(O.pseudo_diffuse_min4(*z["pol"]), z["ed"])
    assert _rel(O.ssim(O.rescale_01(z["ssim_a"]), O.rescale_01(z["ssim_b"]), 5.0), z["ssim"]) < 1e-10
    assert _rel(O.gram_matrix(z["ssim_a"]), z["gram"]) < 1e-12


# ---- GPU: kernels vs fixtures -----------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("tag", CONV_TAGS)
def test_gpu_conv_layers_match_golden(tag):
    from shmgan_b200 import ops
    z = _load("conv_layers")
    N, H, W, Ci, Co, k, s, tr = [int(v) for v in z[tag + "_cfg"]]
    c = ops.Conv("t", k, k, Ci, Co, stride=s, transposed=bool(tr), act=ops.ACT_NONE, bias=True)
    c.w, c.b = z[tag + "_w"].float().cuda(), z[tag + "_b"].float().cuda()
    c.dw, c.db = torch.zeros_like(c.w), torch.zeros(Co, device="cuda")
    x, dy = z[tag + "_x"].float().cuda(), z[tag + "_dy"].float().cuda()
    assert _rel(c.fwd(x, tc=False), z[tag + "_y"]) < 1e-4
    assert _rel(c.dgrad(dy, x.shape, tc=False), z[tag + "_dx"]) < 1e-4
    c.wgrad(x, dy, tc=False)
    assert _rel(c.dw, z[tag + "_dw"]) < 1e-4
    assert _rel(c.db, z[tag + "_dy"].sum(dim=(0, 1, 2))) < 1e-4


@pytest.mark.gpu
def test_gpu_networks_match_golden():
    from shmgan_b200 import nets
    z = _load("nets")
    fs, S = int(z["fs"]), int(z["S"])
    Gp = O.init_params(O.generator_param_specs(fs, True), 1, F, randomize_all=True)       # parameters only (seeded initialiser)
    Dp = O.init_params(O.discriminator_param_specs(S, fs, True), 2, F, randomize_all=True)
    Gn = nets.Generator(fs, True, torch.float32)
    Gn.store.load(Gp)
    feats, _ = Gn.attention(z["mask"].float().cuda())
    assert _rel(Gn.forward(z["x"].float().cuda(), feats), z["g_out"]) < 1e-3
    zero_feats, _ = Gn.attention(torch.zeros_like(z["mask"]).float().cuda())
    Dn = nets.Discriminator(S, fs, True, torch.float32)
    Dn.store.load(Dp)
    attn, _ = Dn.attention(z["mask"].float().cuda())
    rf, cls = Dn.forward(z["img"].float().cuda(), attn)
    assert _rel(rf, z["d_rf"]) < 1e-3 and _rel(cls, z["d_cls"]) < 1e-3


@pytest.mark.gpu
def test_gpu_train_step_matches_golden():
    from shmgan_b200 import model as M
    z = _load("train_step")
    fs, S = 4, 32
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=1, filter_size=fs), dtype="fp32", allow_random_specseg=True).build()
    net.G.net.store.load(O.init_params(O.generator_param_specs(fs, True), 1, F, randomize_all=True))
    net.D.net.store.load(O.init_params(O.discriminator_param_specs(S, fs, True), 2, F, randomize_all=True))
    # pin the mask the fixture used: replace the SpecSeg prediction by the stored mask
    mask = z["mask"].float().cuda()
    net.SpecSeg.net.predict = lambda x, verbose=0: mask
    net.drop_bits, net.TARGET_LABELS = [bool(b) for b in z["bits"]], float(z["T"])
    B = 1
    net.d_noise = torch.zeros((2 * B, S, S, 3), device="cuda")
    net.d_keep = torch.full((2 * B, S // 32, S // 32, fs * 16), 0.8, device="cuda")          # keep * 1/(1-0.2) == 1: dropout off
    net.train_step(*[o.float().cuda() for o in z["origs"]])
    for k in z:
        if k.startswith("loss_"):
            assert getattr(net, k[5:]) == pytest.approx(float(z[k]), rel=2e-3, abs=1e-6), k
    assert _rel(net.gen_Y, z["gen_Y"]) < 1e-3 and _rel(net.gen_rgb, z["gen_rgb"]) < 1e-3


@pytest.mark.gpu
def test_gpu_prep_matches_golden():
    from shmgan_b200 import ops
    z = _load("prep")
    yuv, _ = ops.yuv_standardize(z["img"].float().cuda())
    assert _rel(yuv, z["yuv"]) < 1e-5
    pol = [p.cuda() for p in z["pol"]]
    assert torch.equal(ops.pseudo_diffuse_min4(*pol).cpu(), z["ed"])
