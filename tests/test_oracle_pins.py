"""Second-source pins of the oracle's library semantics that the reference leaves to TensorFlow / Keras (VERDICT r01 "missing" #1):
`tf.image.ssim`, the `rgb_to_yuv` / `yuv_to_rgb` kernels and the Keras Adam update are restated from the published TF 2.8 definitions in
oracle/shmgan_oracle.py; TensorFlow cannot be installed here, so each restatement is checked against an INDEPENDENT implementation
(scipy / numpy / OpenCV / torch.optim / a closed form) that shares no code with it.  The parity of the whole oracle stays "unpinned"
against TF itself; these tests bound the part of that risk that is checkable offline."""
import math

import numpy as np
import pytest
import torch

import oracle as O

F64 = torch.float64


# ---------------------------------------------------------------------------------------------------------------------
# tf.image.ssim (ShmGANwithSSpecSeg.py:759): Wang et al. 2004 with an 11 x 11 Gaussian window (sigma 1.5), VALID, k1 .01, k2 .03
# ---------------------------------------------------------------------------------------------------------------------
def _ssim_independent(a, b, max_val, size=11, sigma=1.5, k1=0.01, k2=0.03):
    """Variance / covariance form of SSIM with a separable window built by scipy.signal.windows.gaussian; numpy float64, NHWC -> [B]."""
    from scipy.ndimage import correlate1d
    from scipy.signal.windows import gaussian
    w = gaussian(size, sigma)
    w = w / w.sum()
    r = size // 2

    def blur(t):                                        # 'valid' part of the separable correlation over H and W
        t = correlate1d(correlate1d(t, w, axis=1, mode="constant"), w, axis=2, mode="constant")
        return t[:, r:-r, r:-r, :]
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    mx, my = blur(a), blur(b)
    vx, vy, cxy = blur(a * a) - mx * mx, blur(b * b) - my * my, blur(a * b) - mx * my
    m = ((2 * mx * my + c1) * (2 * cxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))
    return m.mean(axis=(1, 2)).mean(axis=1)


@pytest.mark.parametrize("max_val", [5.0, 1.0])
def test_ssim_matches_independent_scipy_implementation(max_val):
    g = torch.Generator().manual_seed(3)
    a = torch.rand((2, 40, 48, 3), generator=g, dtype=F64)
    b = (a + 0.15 * torch.randn((2, 40, 48, 3), generator=g, dtype=F64)).clamp(0, 1)
    got = O.ssim(a * max_val, b * max_val, max_val).numpy()
    want = _ssim_independent(a.numpy() * max_val, b.numpy() * max_val, max_val)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-12), (got, want)
    assert np.allclose(O.ssim(a, a, 1.0).numpy(), 1.0, atol=1e-12)           # identical images
    assert np.all(want < 0.99)                                                # ... and the noisy pair is a real test


def test_ssim_gaussian_window_is_the_normalised_sampled_gaussian():
    k = O.shmgan_oracle._gauss_kernel(11, 1.5).numpy()
    x = np.arange(11) - 5.0
    g1 = np.exp(-x * x / (2 * 1.5 ** 2))
    g2 = np.outer(g1, g1)
    assert np.allclose(k, g2 / g2.sum(), rtol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# tf.image.rgb_to_yuv / yuv_to_rgb (:480, :553): analogue BT.601 YUV (W_R .299, W_B .114, U_max .436, V_max .615)
# ---------------------------------------------------------------------------------------------------------------------
def test_yuv_matrices_match_the_bt601_definition_and_invert_each_other():
    wr, wb, umax, vmax = 0.299, 0.114, 0.436, 0.615
    wg = 1.0 - wr - wb
    y = np.array([wr, wg, wb])
    u = umax * (np.array([0.0, 0.0, 1.0]) - y) / (1.0 - wb)
    v = vmax * (np.array([1.0, 0.0, 0.0]) - y) / (1.0 - wr)
    want = np.stack([y, u, v], axis=1)                                        # out = x @ K
    K = np.array(O.RGB2YUV)
    assert np.abs(K - want).max() < 5e-5, np.abs(K - want).max()             # TF's constants are the 8-digit published ones
    Kinv = np.array(O.YUV2RGB)
    assert np.abs(K @ Kinv - np.eye(3)).max() < 2e-6
    # the published inverse: R = Y + 1.13983 V, G = Y - 0.39465 U - 0.58060 V, B = Y + 2.03211 U
    assert np.abs(Kinv - np.array([[1, 1, 1], [0, -0.39465, 2.03211], [1.13983, -0.58060, 0]])).max() < 1e-4


def test_rgb_to_yuv_agrees_with_opencv():
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(4)
    rgb = torch.rand((1, 16, 16, 3), generator=g, dtype=torch.float32)
    got = O.rgb_to_yuv(rgb.double())[0].numpy()
    ref = cv2.cvtColor(rgb[0].numpy(), cv2.COLOR_RGB2YUV)                     # float32: Y, U = .492 (B-Y) + .5, V = .877 (R-Y) + .5
    assert np.abs(got[..., 0] - ref[..., 0]).max() < 1e-5
    assert np.abs(got[..., 1] - (ref[..., 1] - 0.5)).max() < 2e-3             # OpenCV rounds the chroma gains to 3 digits
    assert np.abs(got[..., 2] - (ref[..., 2] - 0.5)).max() < 2e-3
    back = O.yuv_to_rgb(O.rgb_to_yuv(rgb.double()))
    assert float((back - rgb.double()).abs().max()) < 5e-6


# ---------------------------------------------------------------------------------------------------------------------
# Keras Adam + ExponentialDecay (:169-175): epsilon OUTSIDE the bias-corrected square root
# ---------------------------------------------------------------------------------------------------------------------
def test_keras_adam_matches_closed_form_for_a_constant_gradient():
    """With a constant gradient g: m_t = g (1 - b1^t), v_t = g^2 (1 - b2^t), so every step is -lr_t m_t / (sqrt(v_t) + eps)."""
    b1, b2, eps, lr0 = 0.5, 0.99, 1e-7, 2e-5
    g = torch.tensor([0.3, -1.0, 1e-6, 0.0], dtype=F64)
    P, m, v = {"w": torch.zeros(4, dtype=F64)}, {"w": torch.zeros(4, dtype=F64)}, {"w": torch.zeros(4, dtype=F64)}
    want = torch.zeros(4, dtype=F64)
    for step in range(3):
        P, m, v = O.keras_adam_update(P, {"w": g}, m, v, step, lr0, b1, b2, eps)
        t = step + 1
        lr = lr0 * 0.95 ** (step / 10000.0)
        lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        want = want - lr_t * (g * (1 - b1 ** t)) / ((g * g * (1 - b2 ** t)).sqrt() + eps)
        assert torch.allclose(P["w"], want, rtol=1e-12, atol=1e-18), (step, P["w"], want)


def test_keras_adam_matches_torch_adam_with_the_epsilon_moved():
    """torch.optim.Adam puts eps inside: lr m_hat / (sqrt(v_hat) + eps) = lr_t m / (sqrt(v) + eps sqrt(1 - b2^t)); running it with
    eps_t = eps / sqrt(1 - b2^t) and the decayed lr per step is the Keras update -- an independent implementation of the same formula."""
    b1, b2, eps, lr0 = 0.5, 0.99, 1e-7, 2e-5
    gen = torch.Generator().manual_seed(5)
    w0 = torch.randn(64, generator=gen, dtype=F64)
    grads = [torch.randn(64, generator=gen, dtype=F64) * s for s in (1.0, 1e-3, 0.2)]
    P, m, v = {"w": w0.clone()}, {"w": torch.zeros(64, dtype=F64)}, {"w": torch.zeros(64, dtype=F64)}
    wt = w0.clone().requires_grad_()
    opt = torch.optim.Adam([wt], lr=lr0, betas=(b1, b2), eps=eps)
    for step, g in enumerate(grads):
        P, m, v = O.keras_adam_update(P, {"w": g}, m, v, step, lr0, b1, b2, eps)
        t = step + 1
        for grp in opt.param_groups:
            grp["lr"] = O.keras_adam_lr(step, lr0)
            grp["eps"] = eps / math.sqrt(1 - b2 ** t)
        wt.grad = g.clone()
        opt.step()
        assert torch.allclose(P["w"], wt.detach(), rtol=1e-11, atol=1e-16), step


# ---------------------------------------------------------------------------------------------------------------------
# tfa InstanceNormalization / Keras BatchNormalization(eval) / softmax CE against torch's own functional forms
# ---------------------------------------------------------------------------------------------------------------------
def test_instance_norm_and_softmax_ce_match_torch_functional():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(6)
    x = torch.randn((2, 9, 7, 5), generator=g, dtype=F64)
    gamma, beta = torch.rand(5, generator=g, dtype=F64) + 0.5, torch.randn(5, generator=g, dtype=F64)
    want = F.instance_norm(x.permute(0, 3, 1, 2), weight=gamma, bias=beta, eps=1e-6).permute(0, 2, 3, 1)
    assert torch.allclose(O.instance_norm(x, gamma, beta), want, rtol=1e-10, atol=1e-12)
    logits = torch.randn((4, 5), generator=g, dtype=F64)
    lab = torch.tensor([[0, 0, 0, 0, 0.93]] * 4, dtype=F64)
    want = -(lab * F.log_softmax(logits, dim=1)).sum(1)
    assert torch.allclose(O.softmax_ce(lab, logits), want, rtol=1e-12)
    hard = F.cross_entropy(logits, torch.tensor([1, 1, 1, 1]), reduction="none")
    assert torch.allclose(O.softmax_ce(torch.eye(5, dtype=F64)[1][None].expand(4, 5), logits), hard, rtol=1e-12)


def test_conv_transpose_same_is_the_input_gradient_of_the_strided_same_conv():
    """Keras Conv2DTranspose(k, s=2, 'same') == conv2d_backprop_input of the matching SAME conv (TF's own definition): checked through
    torch autograd of the oracle's conv2d_same, which shares no code with conv2d_transpose_same."""
    g = torch.Generator().manual_seed(7)
    for k in (3, 2):
        x = torch.randn((1, 6, 5, 4), generator=g, dtype=F64)                 # the SMALL image (Conv2DTranspose input), Cin = 4
        w = torch.randn((k, k, 3, 4), generator=g, dtype=F64)                 # (kh, kw, Cout, Cin)
        big = torch.zeros((1, 12, 10, 3), dtype=F64, requires_grad=True)
        y = O.conv2d_same(big, w, None, 2)                                    # forward conv big -> small with kernel (kh,kw,Cin=3,Cout=4)
        gx, = torch.autograd.grad((y * x).sum(), big)
        assert torch.allclose(O.conv2d_transpose_same(x, w, None, 2), gx, rtol=1e-12, atol=1e-12), k
