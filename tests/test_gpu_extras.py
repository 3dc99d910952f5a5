"""GPU parity of the SURVEY 8(f) rows through the C ABI: loader contract (bit-exact), degree of polarisation (bit-exact), test-time
metrics (float tolerance written per metric), checkpoint round trip of the real networks and the loader feeding train_step."""
import numpy as np
import pytest
import torch

from oracle import extras_oracle as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("src,dst,flip", [((256, 256), 256, False), ((256, 256), 256, True), ((100, 140), 256, True), ((300, 400), 256, False),
                                           ((37, 53), 64, True), ((1024, 1024), 512, False), ((5, 5), 32, False)])
def test_loader_kernel_bit_exact(src, dst, flip):
    from shmgan_b200 import ops
    rng = np.random.default_rng(src[0] * 7 + dst)
    img = rng.integers(0, 256, size=(3, src[0], src[1], 3), dtype=np.uint8)
    want = E.load_images(img, dst, random_flip=not flip)
    got = ops.load_u8_images(torch.from_numpy(img).cuda(), dst, flip).cpu().numpy()
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)                              # same float32 order of operations as TF's ResizeBilinear + x/255


def test_loader_empty_batch():
    from shmgan_b200 import ops
    out = ops.load_u8_images(torch.zeros((0, 8, 8, 3), dtype=torch.uint8, device="cuda"), 16, False)
    assert out.shape == (0, 16, 16, 3)


def test_dop_bit_exact_and_zero_intensity():
    from shmgan_b200 import ops
    rng = np.random.default_rng(11)
    planes = [rng.random((4, 64, 64, 1)).astype(np.float32) for _ in range(4)]
    planes[0][0, :8] = 0.0; planes[2][0, :8] = 0.0                 # S0 = 0 rows -> divide_no_nan -> 0
    want_d, want_a = E.dop(*planes)
    got_d, got_a = ops.dop(*[torch.from_numpy(p).cuda() for p in planes], want_angle=True)
    assert np.array_equal(got_d.cpu().numpy(), want_d)
    assert float(got_d[0, :8].abs().max()) == 0.0
    assert np.abs(got_a.cpu().numpy() - want_a).max() < 1e-6      # atan2f is not correctly rounded: float tolerance


@pytest.mark.parametrize("n,s", [(1, 256), (3, 64)])
def test_image_metrics_parity(n, s):
    from shmgan_b200 import metrics
    rng = np.random.default_rng(n * 100 + s)
    tgt = rng.random((n, s, s, 3)).astype(np.float32)
    gen = np.clip(tgt + 0.05 * rng.standard_normal(tgt.shape), 0, 1).astype(np.float32)
    got = metrics.image_metrics(torch.from_numpy(gen).cuda(), torch.from_numpy(tgt).cuda())
    assert got["mse"] == pytest.approx(E.mse(gen, tgt), rel=1e-6)
    assert np.allclose(got["psnr"], E.psnr(gen, tgt, 1.0), atol=1e-4)
    for i in range(n):                                             # rescale_01 is per image here (Q5); the reference tests one image at a time
        w = E.image_metrics(gen[i:i + 1], tgt[i:i + 1])
        assert got["ssim"][i] == pytest.approx(float(w["ssim"][0]), abs=2e-4)
        assert got["delE76"][i] == pytest.approx(w["delE76"], rel=1e-3)     # float32 gamma / cube root against the float64 oracle
        assert got["delE94"][i] == pytest.approx(w["delE94"], rel=1e-3)
    same = metrics.image_metrics(torch.from_numpy(tgt).cuda(), torch.from_numpy(tgt).cuda())
    assert same["mse"] == 0.0 and same["delE76"] == [0.0] * n and all(abs(v - 1.0) < 1e-4 for v in same["ssim"])


def _net(seed_shift=0):
    from shmgan_b200 import model as M
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=64, batch_size=2, filter_size=8), dtype="bf16", allow_random_specseg=True).build()
    if seed_shift:
        net.G.net.store.init(100 + seed_shift); net.D.net.store.init(200 + seed_shift)
    net.drop_bits, net.TARGET_LABELS, net.noise_seed = [True, False, False, True, False], 0.9, 5
    return net


def test_checkpoint_roundtrip_resumes_training(tmp_path):
    from shmgan_b200.checkpoint import Checkpoint, CheckpointManager
    g = torch.Generator().manual_seed(0)
    batch = [torch.rand((2, 64, 64, 3), generator=g).cuda() for _ in range(5)]
    a = _net()
    a.drop_bits = None                                             # let the host RNG draw the five bits: its state must survive the resume
    a.train_step(*batch)                                           # Adam moments and step counters are now non-trivial
    mgr = CheckpointManager(Checkpoint(generator=a.G, discriminator=a.D, specseg=a.SpecSeg, host=a), str(tmp_path), max_to_keep=3)
    path = mgr.save()
    b = _net(seed_shift=1)
    b.drop_bits = None
    b.SpecSeg.net.store.init(300)
    assert not torch.equal(a.G.net.store.flat, b.G.net.store.flat)
    Checkpoint(generator=b.G, discriminator=b.D, specseg=b.SpecSeg, host=b).restore(
        CheckpointManager(Checkpoint(), str(tmp_path)).latest_checkpoint).assert_consumed()
    # SpecSeg travels with the checkpoint (the reference re-reads specsegv3_chkpt.h5 instead, :931) and counts as loaded weights
    assert torch.equal(a.SpecSeg.net.store.flat, b.SpecSeg.net.store.flat) and b.SpecSeg.loaded
    # host-side state of a bit-reproducible resume: Philox step counter, drop-bit RNG, D call counter, running standardisation scale
    assert b.step_count == a.step_count == 1 and b._rng.getstate() == a._rng.getstate() and b.D.calls == a.D.calls
    assert b.stddev_arr.state() == a.stddev_arr.state() and b.noise_seed == a.noise_seed
    assert path == mgr.latest_checkpoint
    for x, y in ((a.G, b.G), (a.D, b.D)):
        sx, sy = x.net.store, y.net.store
        assert torch.equal(sx.flat, sy.flat) and torch.equal(sx.m, sy.m) and torch.equal(sx.v, sy.v) and sx.step == sy.step == 1
    for net in (a, b):
        net.train_step(*batch)                                     # bits, noise and dropout now come from the restored RNG state
    assert a.last_drop_bits == b.last_drop_bits
    assert a.total_Generator_loss == pytest.approx(b.total_Generator_loss, rel=1e-5)
    assert a.total_Discriminator_loss == pytest.approx(b.total_Discriminator_loss, rel=1e-5)
    # fp32 atomics order only: the gradients agree to ~1e-6 relative, so Adam's normalised step (~2e-5 per weight) agrees except where a
    # gradient's sign sits inside that noise
    d = (a.G.net.store.flat - b.G.net.store.flat).abs()
    assert float(d.max()) < 5e-5 and float((d > 1e-6).float().mean()) < 1e-3


def test_loader_feeds_train_step():
    from shmgan_b200 import loader
    rng = np.random.default_rng(21)
    pol = [[rng.integers(0, 256, size=(80, 96, 3), dtype=np.uint8) for _ in range(4)] for _ in range(4)]
    ed = [np.minimum(np.minimum(pol[0][i], pol[1][i]), np.minimum(pol[2][i], pol[3][i])) for i in range(4)]
    ld = loader.PolarimetricLoader(pol + [ed], image_size=64, batch_size=2, random_flip=False, repeat=1)
    batches = list(ld)
    assert len(batches) == len(ld) == 2 and all(len(b) == 5 for b in batches)
    want = E.load_images(np.stack(pol[2][2:4]), 64, random_flip=False)
    assert np.array_equal(batches[1][2].cpu().numpy(), want)
    net = _net()
    net.train_step(*batches[0])
    assert np.isfinite(net.total_Generator_loss) and np.isfinite(net.total_Discriminator_loss)


def test_loader_estimates_diffuse_on_device():
    from shmgan_b200 import loader
    rng = np.random.default_rng(22)
    pol = [[rng.integers(0, 256, size=(70, 90, 3), dtype=np.uint8) for _ in range(3)] for _ in range(4)]
    ld = loader.PolarimetricLoader(pol, image_size=64, batch_size=3, random_flip=True, est_diffuse=True)
    (b,) = list(ld)
    assert len(b) == 5
    ed = np.minimum(np.minimum(np.stack(pol[0]), np.stack(pol[1])), np.minimum(np.stack(pol[2]), np.stack(pol[3])))
    assert np.array_equal(b[4].cpu().numpy(), E.load_images(ed, 64, random_flip=True))
