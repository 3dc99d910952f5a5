"""CPU tests of the SURVEY 8(f) rows: the numpy oracle (oracle/extras_oracle.py) pinned against independent implementations in this
image (torch's half-pixel bilinear resize, OpenCV's float RGB->Lab, hand-computed Delta-E / DoP / PSNR values), and the host-side
logic of the checkpoint manager and the loader (no kernels)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import extras_oracle as E


@pytest.mark.parametrize("src,dst", [((37, 53), (64, 64)), ((300, 400), (256, 256)), ((64, 64), (64, 64)), ((5, 7), (32, 16)), ((129, 65), (64, 32))])
def test_resize_matches_torch_half_pixel_bilinear(src, dst):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(2, src[0], src[1], 3), dtype=np.uint8)
    got = E.resize_bilinear_tf2(img, dst[0], dst[1])
    want = F.interpolate(torch.from_numpy(img).permute(0, 3, 1, 2).double(), size=dst, mode="bilinear", align_corners=False,
                         antialias=False).permute(0, 2, 3, 1).numpy()
    assert got.dtype == np.float32 and got.shape == (2, dst[0], dst[1], 3)
    assert np.abs(got - want).max() < 2e-3                        # float32 lerp order against a float64 evaluation of 0..255 values


def test_resize_identity_and_constant():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(1, 32, 48, 3), dtype=np.uint8)
    assert np.array_equal(E.resize_bilinear_tf2(img, 32, 48), img.astype(np.float32))        # scale 1 -> lerp 0 -> exact
    const = np.full((1, 9, 11, 3), 77, np.uint8)
    assert np.array_equal(E.resize_bilinear_tf2(const, 40, 23), np.full((1, 40, 23, 3), 77, np.float32))


def test_load_images_scale_and_flip_polarity():
    img = np.arange(2 * 4 * 4 * 3, dtype=np.uint8).reshape(2, 4, 4, 3)
    kept = E.load_images(img, 4, random_flip=True)                 # datasetLoader.py:61: random_flip True -> NOT flipped
    flipped = E.load_images(img, 4, random_flip=False)
    assert np.array_equal(kept, img.astype(np.float32) / np.float32(255.0))
    assert np.array_equal(flipped, kept[:, ::-1])
    assert kept.max() <= 1.0 and kept.dtype == np.float32


def test_dop_hand_values():
    i0, i45, i90, i135 = (np.array(v, np.float32) for v in ([1.0, 0.0, 0.6, 0.5], [0.5, 0.0, 0.9, 0.5], [0.0, 0.0, 0.2, 0.5], [0.5, 0.0, 0.1, 0.5]))
    d, a = E.dop(i0, i45, i90, i135)
    assert d[0] == 1.0 and d[1] == 0.0                            # fully polarised; S0 = 0 -> divide_no_nan -> 0
    assert d[2] == pytest.approx(math.sqrt(0.4 ** 2 + 0.8 ** 2) / 0.8, rel=1e-6)
    assert d[3] == 0.0 and a[2] == pytest.approx(0.5 * math.atan2(0.8, 0.4), rel=1e-6)


def test_rgb_to_lab_against_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    rgb = rng.random((64, 64, 3)).astype(np.float32)
    want = cv2.cvtColor(rgb, cv2.COLOR_RGB2Lab)                    # float input: L in [0,100], a/b in about [-127,127], D65
    got = E.rgb_to_lab(rgb)
    # OpenCV's float path interpolates tabulated gamma / cube-root curves (outputs land on a 1/64 grid): a convention pin, not an ulp pin
    assert np.abs(got - want).max() < 0.6 and np.abs(got - want).mean() < 0.1
    white = E.rgb_to_lab(np.ones((1, 3)))
    assert white[0, 0] == pytest.approx(100.0, abs=1e-2) and abs(white[0, 1]) < 1e-2 and abs(white[0, 2]) < 2e-2
    assert np.allclose(E.rgb_to_lab(np.zeros((1, 3))), 0.0, atol=1e-9)


def test_delta_e_hand_values():
    ref = np.array([[50.0, 10.0, 0.0]])
    assert E.delta_e76(ref, ref)[0] == 0.0 and E.delta_e94(ref, ref)[0] == 0.0
    assert E.delta_e76(ref, np.array([[53.0, 14.0, 0.0]]))[0] == pytest.approx(5.0)
    assert E.delta_e94(ref, np.array([[60.0, 10.0, 0.0]]))[0] == pytest.approx(10.0)                     # lightness only, S_L = 1
    assert E.delta_e94(ref, np.array([[50.0, 20.0, 0.0]]))[0] == pytest.approx(10.0 / 1.45)              # chroma only, S_C = 1 + 0.045 * 10
    assert E.delta_e94(ref, np.array([[50.0, 0.0, 10.0]]))[0] == pytest.approx(math.sqrt(200.0) / 1.15)  # hue only, S_H = 1 + 0.015 * 10
    # asymmetric: the first argument is the reference colour
    assert E.delta_e94(np.array([[50.0, 20.0, 0.0]]), ref)[0] == pytest.approx(10.0 / 1.9)


def test_psnr_mse_hand_values():
    a = np.zeros((2, 4, 4, 3)); b = np.zeros((2, 4, 4, 3))
    b[0] += 0.1; b[1] += 0.01
    assert E.mse(a, b) == pytest.approx((0.01 + 0.0001) / 2)
    p = E.psnr(a, b, 1.0)
    assert p[0] == pytest.approx(20.0) and p[1] == pytest.approx(40.0)
    m = E.image_metrics(np.random.default_rng(3).random((1, 32, 32, 3)), np.random.default_rng(4).random((1, 32, 32, 3)))
    assert set(m) == {"mse", "ssim", "psnr", "delE76", "delE94"} and 0 < m["delE94"] < m["delE76"]
    same = E.image_metrics(np.random.default_rng(3).random((1, 32, 32, 3)), np.random.default_rng(3).random((1, 32, 32, 3)))
    assert same["ssim"][0] == pytest.approx(1.0) and same["delE76"] == 0.0


# ---- host logic: checkpoint manager with a stand-in parameter store (same attributes as nets.ParamStore) -----------------------
class _Store:
    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)
        self.offsets = {"conv2d.w": (0, 24, (2, 3, 4)), "conv2d.b": (24, 4, (4,)), "in.beta": (28, 4, (4,))}
        self.n_train = 28
        self.flat = torch.randn(32, generator=g)
        self.m, self.v = torch.randn(28, generator=g), torch.rand(28, generator=g)
        self.gviews = {"conv2d.w": None, "conv2d.b": None}
        self.step = seed

    def export(self):
        return {k: self.flat[o:o + n].view(s).clone() for k, (o, n, s) in self.offsets.items()}

    def load(self, named):
        for k, t in named.items():
            o, n, _ = self.offsets[k]
            self.flat[o:o + n] = t.reshape(-1)


class _Net:
    def __init__(self, seed):
        self.store = _Store(seed)


def test_checkpoint_manager_roundtrip_and_pruning(tmp_path):
    from shmgan_b200.checkpoint import Checkpoint, CheckpointManager
    G, D = _Net(1), _Net(2)
    mgr = CheckpointManager(Checkpoint(generator=G, discriminator=D), str(tmp_path), max_to_keep=3)
    assert mgr.latest_checkpoint is None
    paths = []
    for i in range(5):
        G.store.step = 10 + i
        paths.append(mgr.save())
    assert [os.path.basename(p) for p in mgr.checkpoints] == ["ckpt-3.npz", "ckpt-4.npz", "ckpt-5.npz"]
    assert not os.path.exists(paths[0]) and os.path.exists(paths[4])
    G2, D2 = _Net(7), _Net(8)
    mgr2 = CheckpointManager(Checkpoint(generator=G2, discriminator=D2), str(tmp_path), max_to_keep=3)   # a fresh process finds the index
    assert mgr2.latest_checkpoint == paths[4]
    st = Checkpoint(generator=G2, discriminator=D2).restore(mgr2.latest_checkpoint).expect_partial()
    st.assert_consumed()
    for a, b in ((G, G2), (D, D2)):
        assert torch.equal(a.store.flat[:32], b.store.flat[:32]) and torch.equal(a.store.m, b.store.m) and torch.equal(a.store.v, b.store.v)
    assert G2.store.step == 14 and D2.store.step == 2
    assert mgr2.save().endswith("ckpt-6.npz")
    Checkpoint(generator=G2).restore(None).expect_partial()        # no checkpoint yet: no-op, like tf.train.Checkpoint.restore(None)
    with np.load(paths[4]) as z:                                    # reference variable names / layouts are the npz keys
        assert "generator/conv2d.w" in z.files and z["generator/adam_m/conv2d.w"].shape == (2, 3, 4) and "generator/adam_m/in.beta" not in z.files


def test_loader_lists_and_decodes_in_sorted_order(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    from shmgan_b200 import loader
    rng = np.random.default_rng(5)
    imgs = {}
    for name in ("b_02.png", "a_10.png", "a_02.png", "notes.txt"):
        if name.endswith(".png"):
            imgs[name] = rng.integers(0, 256, size=(6, 5, 3), dtype=np.uint8)
            Image.fromarray(imgs[name]).save(tmp_path / name)
        else:
            (tmp_path / name).write_text("x")
    files = loader.list_folder(str(tmp_path))
    assert [os.path.basename(f) for f in files] == ["a_02.png", "a_10.png", "b_02.png"]
    assert np.array_equal(loader.decode_rgb_u8(files[1]), imgs["a_10.png"])
    ld = loader.PolarimetricLoader([[imgs["a_02.png"]] * 3] * 5, image_size=32, batch_size=2, random_flip=True, repeat=2)
    assert ld.length_dataset == 3 and len(ld) == 4 and ld.flip is False
    with pytest.raises(AssertionError):
        loader.PolarimetricLoader([[imgs["a_02.png"]] * 3] * 4 + [[imgs["a_02.png"]] * 2], image_size=32)
