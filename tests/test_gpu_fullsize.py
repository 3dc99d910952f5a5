"""GPU tests at BASELINE.json's FULL sizes (configs[1]: batch 16 at 256 x 256 -> the 5B = 80-image generator pass), where the CPU oracle
would take minutes: size-independent properties instead of element-wise comparison.

  * convolutions: translation equivariance (bit-exact away from the border), linearity, agreement of the tensor-core kernels with the
    exact-fp32 SIMT kernels on a sampled sub-batch, weight-gradient additivity over the batch;
  * instance norm: the output's per-(n, c) moments are (beta, gamma^2); the backward's output is orthogonal to 1 and to x_hat;
    the cp.async-pipelined kernels agree with the register-staged ones;
  * train step: per-sample results do not depend on the batch they are computed in (the data-parallel premise), and two
    runs of the same step agree.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

N, S = 80, 256          # the 5B cyclic generator pass of configs[1]


def _conv(cin, cout, seed, k=3, stride=1, transposed=False, act=0, bias=True):
    from shmgan_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    c = ops.Conv("t", k, k, cin, cout, stride=stride, transposed=transposed, act=act, bias=bias)
    wshape = (k, k, cout, cin) if transposed else (k, k, cin, cout)
    c.w = (torch.randn(wshape, device="cuda", generator=g) * 0.05).bfloat16().float()
    c.b = torch.randn(cout, device="cuda", generator=g) * 0.1 if bias else None
    c.dw = torch.zeros_like(c.w)
    c.db = torch.zeros(cout, device="cuda") if bias else None
    return c


def _x(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g).bfloat16()


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 256), (128, 128, 128), (512, 256, 64)])
def test_conv_full_size_translation_equivariance_bit_exact(cin, cout, hw):
    """A stride-1 SAME convolution commutes with translation: shifting the input by 32 rows and 8 columns shifts the output, and every output
    pixel sees the same products in the same order, so the interior is bit-identical (halo / big-halo kernels, 80 images)."""
    c = _conv(cin, cout, 1, act=1)
    x = _x((N, hw, hw, cin), 2)
    y = c.fwd(x, tc=True, version=1)
    xs = torch.zeros_like(x)
    xs[:, 32:, 8:] = x[:, :-32, :-8]
    ys = c.fwd(xs, tc=True, version=1)
    # interior only: rows / columns whose 3 x 3 window touches neither image border in either run
    assert torch.equal(ys[:, 34:-2, 10:-2], y[:, 2:-34, 2:-10])
    assert float(y.float().abs().mean()) > 1e-3


def test_conv_full_size_linearity_and_simt_agreement():
    """conv(x1 + 2 x2) = conv(x1) + 2 conv(x2) (no bias, no activation) within bf16 output rounding, at 80 x 256 x 256 x 64; and the
    tensor-core result of a sampled sub-batch agrees with the exact-fp32 SIMT kernel on the same bf16 inputs."""
    from _util import rel_err
    c = _conv(64, 64, 3, act=0, bias=False)
    x1, x2 = _x((N, S, S, 64), 4), _x((N, S, S, 64), 5)
    xs = (x1.float() + 2.0 * x2.float()).bfloat16()
    lhs = c.fwd(xs, tc=True, version=1).float()
    rhs = c.fwd(x1, tc=True, version=1).float() + 2.0 * c.fwd(x2, tc=True, version=1).float()
    # inputs of the left side are re-rounded to bf16 (2^-9 relative per element, random sign) and all three outputs are bf16
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 2e-2
    sub = x1[37:39].contiguous()
    ref = c.fwd(sub.float(), tc=False)
    assert rel_err(c.fwd(sub, tc=True, version=1), ref.double()) < 1e-2


def test_wgrad_full_size_is_additive_over_the_batch():
    """dW(all 80 images) = dW(first 48) + dW(last 32): the split-K weight-gradient kernels accumulate (`dw +=`) and every image is visited once."""
    c = _conv(64, 64, 6)
    x, dy = _x((N, S, S, 64), 7), _x((N, S, S, 64), 8)
    c.wgrad(x, dy, tc=True)
    whole = c.dw.clone()
    c.dw.zero_()
    c.wgrad(x[:48].contiguous(), dy[:48].contiguous(), tc=True)
    c.wgrad(x[48:].contiguous(), dy[48:].contiguous(), tc=True)
    assert float((c.dw - whole).abs().max() / whole.abs().max()) < 1e-4        # fp32 atomics in a different order


@pytest.mark.parametrize("shape", [(N, S, S, 64), (N, 128, 128, 128), (N, 32, 32, 512)])
def test_instance_norm_full_size_moments_and_backward_orthogonality(shape):
    from shmgan_b200 import ops
    n, h, w, c = shape
    x = (_x(shape, 9).float() * 1.7 + 0.4).bfloat16()
    g = torch.Generator(device="cuda").manual_seed(10)
    gamma = torch.rand(c, device="cuda", generator=g) + 0.5
    beta = torch.randn(c, device="cuda", generator=g) * 0.2
    sums = ops.inorm_stats(x)
    y, pooled = ops.inorm_apply(x, sums, gamma, beta, pooled=True)
    yf = y.float()
    mean, var = yf.mean(dim=(1, 2)), yf.var(dim=(1, 2), unbiased=False)
    assert float((mean - beta).abs().max()) < 5e-3                               # bf16 output rounding averages out over >= 1024 pixels
    assert float((var / (gamma * gamma) - 1.0).abs().max()) < 1e-2
    want_pool = yf.view(n, h // 2, 2, w // 2, 2, c).mean(dim=(2, 4))
    assert float((pooled.float() - want_pool).abs().max()) < 2e-2
    # backward with act = identity: sum_p dx = 0 and sum_p dx * x_hat = 0 for every (n, c)
    dy = _x(shape, 11)
    dx = ops.inorm_bwd(x, sums, gamma, dy, None, act=0).float()
    xf = x.float()
    xhat = (xf - xf.mean(dim=(1, 2), keepdim=True)) / xf.var(dim=(1, 2), unbiased=False, keepdim=True).add(1e-6).sqrt()
    scale = float(dx.abs().mean()) * h * w
    assert float(dx.sum(dim=(1, 2)).abs().max()) / scale < 2e-3
    assert float((dx * xhat).sum(dim=(1, 2)).abs().max()) / scale < 2e-3


def test_pipelined_norm_kernels_match_register_staged_full_size():
    from shmgan_b200 import ops
    from shmgan_b200._lib import call
    shape = (N, S, S, 64)
    x, dy, dyp = _x(shape, 12), _x(shape, 13), _x((N, S // 2, S // 2, 64), 14)
    gamma = torch.ones(64, device="cuda"); beta = torch.zeros(64, device="cuda")

    def run():
        sums = ops.inorm_stats(x)
        y, p = ops.inorm_apply(x, sums, gamma, beta, pooled=True)
        db = torch.zeros(64, device="cuda")
        dx = ops.inorm_bwd(x, sums, gamma, dy, dyp, dbias=db)
        return sums.clone(), y, p, dx, db
    try:
        call("shm_norm_tune", 1, 0, 0)
        ref = run()
    finally:
        call("shm_norm_tune", 0, 0, 0)
    got = run()
    assert float((got[0] - ref[0]).abs().max() / ref[0].abs().max()) < 1e-6      # fp32 partial sums of different lengths
    for a, b in zip(got[1:4], ref[1:4]):
        d = (a.float() - b.float()).abs().max() / b.float().abs().max()
        assert float(d) < 4e-3                                                    # at most one bf16 ulp
    assert float((got[4] - ref[4]).abs().max() / ref[4].abs().max()) < 1e-3


def test_train_step_full_size_batch_independence_and_repeatability():
    """configs[1] shape (B = 16 at 256 x 256, bf16): sample i's generator output is the same whether it is computed in the batch of 16
    or in a batch of 8 (instance norm / standardisation / rescale are per sample), and the batch-mean losses of the two halves average
    to the loss of the whole batch; a second run of the same step (fresh networks, same seeds) reproduces the losses."""
    from shmgan_b200 import model as M

    def fresh(B):
        net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="bf16", allow_random_specseg=True).build()
        net.drop_bits, net.TARGET_LABELS, net.noise_seed = [True, False, True, False, False], 0.9, 3
        return net
    g = torch.Generator().manual_seed(15)
    pol = [torch.rand((16, S, S, 3), generator=g).cuda() for _ in range(4)]
    ed = torch.minimum(torch.minimum(pol[0], pol[1]), torch.minimum(pol[2], pol[3]))
    batch = pol + [ed]
    a = fresh(16)
    a.train_step(*batch)
    whole_Y, whole_G, whole_L1 = a.gen_Y.clone(), a.total_Generator_loss, a.L1_loss_Gen
    del a
    torch.cuda.empty_cache()
    b = fresh(16)
    b.train_step(*batch)
    assert b.total_Generator_loss == pytest.approx(whole_G, rel=1e-4) and b.L1_loss_Gen == pytest.approx(whole_L1, rel=1e-4)
    assert float((b.gen_Y - whole_Y).abs().max()) < 1e-2 * float(whole_Y.abs().max())
    del b
    torch.cuda.empty_cache()
    l1 = []
    for lo in (0, 8):
        h = fresh(8)
        h.train_step(*[t[lo:lo + 8].contiguous() for t in batch])
        d = (h.gen_Y - whole_Y[lo:lo + 8]).abs().max() / whole_Y.abs().max()
        assert float(d) < 1e-2                                                    # bf16 activations; statistics summed in another order
        l1.append(h.L1_loss_Gen)
        del h
        torch.cuda.empty_cache()
    assert 0.5 * (l1[0] + l1[1]) == pytest.approx(whole_L1, rel=2e-3)
