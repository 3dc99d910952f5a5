"""Known-answer tests that pin the CPU oracle against the only numerical ground truth the reference
ships: the Keras summaries (Generator_summary.txt, Discriminator_summary.txt, SpecSeg_summary.txt, all
at 128x128) and hand-checkable padding impulses."""
import math

import pytest
import torch

import oracle as O

F64 = torch.float64


def test_param_counts_match_reference_summaries():
    # Generator_summary.txt:621  "Total params: 18,525,569" (as-written graph: no attention convs)
    assert O.count_params(O.generator_param_specs(64, live_mask=False)) == 18_525_569
    # Discriminator_summary.txt:179 "Total params: 6,359,744" @128x128
    assert O.count_params(O.discriminator_param_specs(128, 64, live_mask=False)) == 6_359_744
    # SpecSeg_summary.txt:118-120: 1,942,801 total / 1,941,809 trainable / 992 non-trainable
    s = O.specseg_param_specs()
    assert O.count_params(s) == 1_942_801
    assert O.count_params(s, trainable_only=True) == 1_941_809
    # live-mask mode adds the 8 attention convs: 3 142 080 weights + 1920 biases
    assert O.count_params(O.generator_param_specs(64, True)) == 18_525_569 + 3_142_080 + 1920


def test_per_layer_param_counts_generator():
    # Generator_summary.txt:7,39,73,105,139,171,205,237,271,303,335,342,374,406,413,445
    want = {"enc1a": 5824, "enc1b": 36928, "enc2a": 73856, "enc2b": 147584, "enc3a": 295168,
            "enc3b": 590080, "enc4a": 1180160, "enc4b": 2359808, "bott1": 262656, "bott2": 262656,
            "up1T": 2359808, "dec1a": 4719104, "dec1b": 2359808, "up2T": 1179904, "dec2a": 1179904,
            "dec2b": 590080}
    got = {}
    for name, shape, kind in O.generator_param_specs(64, False):
        if "in_" in name:
            continue
        got[name.split(".")[0]] = got.get(name.split(".")[0], 0) + math.prod(shape)
    for k, v in want.items():
        assert got[k] == v, k


def test_layer_shapes_at_128():
    # Generator_summary.txt:5-623: input (128,128,10) -> output (128,128,1); skips 128/64/32/16; bottleneck 8x8x512
    p = O.init_params(O.generator_param_specs(64, False), 1, torch.float32)
    y, inter = O.generator_forward(p, torch.rand(1, 128, 128, 10), None, return_intermediates=True)
    assert tuple(y.shape) == (1, 128, 128, 1)
    assert [tuple(s.shape[1:]) for s in inter["skips"]] == [(128, 128, 64), (64, 64, 128), (32, 32, 256), (16, 16, 512)]
    # Discriminator_summary.txt: (64,64,64) ... (4,4,1024) -> rf (4,4,1), cls (5)
    d = O.init_params(O.discriminator_param_specs(128, 64, False), 2, torch.float32)
    rf, cls = O.discriminator_forward(d, torch.rand(1, 128, 128, 3))
    assert tuple(rf.shape) == (1, 4, 4, 1) and tuple(cls.shape) == (1, 5)
    s = O.init_params(O.specseg_param_specs(), 3, torch.float32)
    assert tuple(O.specseg_forward(s, torch.rand(1, 128, 128, 1)).shape) == (1, 128, 128, 1)


def test_same_padding_rule():
    assert O.tf_same_pad(256, 3, 1) == (256, 1, 1)
    assert O.tf_same_pad(256, 3, 2) == (128, 0, 1)      # stride 2, even input: pad goes at the END
    assert O.tf_same_pad(256, 1, 1) == (256, 0, 0)
    assert O.tf_same_pad(7, 3, 2) == (4, 1, 1)


def test_stride2_conv_impulse_alignment():
    # y[o] = sum_k x[2o + k] w[k] (pad_before = 0): impulse at x[2,2] hits y[1,1] via w[0,0], y[0,0] via w[2,2]
    x = torch.zeros(1, 8, 8, 1, dtype=F64)
    x[0, 2, 2, 0] = 1.0
    w = torch.arange(9, dtype=F64).reshape(3, 3, 1, 1) + 1
    y = O.conv2d_same(x, w, None, 2)[0, :, :, 0]
    assert y[1, 1] == w[0, 0, 0, 0] and y[0, 0] == w[2, 2, 0, 0] and y[0, 1] == w[2, 0, 0, 0]
    assert y.sum() == w[0, 0] .sum()+ w[2, 2].sum() + w[2, 0].sum() + w[0, 2].sum()


def test_transposed_conv_impulse_alignment():
    # out[p] = sum_{2o+k=p} x[o] w[k]: impulse at o=(1,1) lands on p = 2..4 with w[k] at p=2+k, cropped to 2*in
    x = torch.zeros(1, 4, 4, 1, dtype=F64)
    x[0, 1, 1, 0] = 1.0
    w = (torch.arange(9, dtype=F64).reshape(3, 3, 1, 1) + 1)
    y = O.conv2d_transpose_same(x, w, None, 2)[0, :, :, 0]
    assert tuple(y.shape) == (8, 8)
    assert torch.equal(y[2:5, 2:5], w[:, :, 0, 0])
    assert y.sum() == w.sum()
    # last input pixel: the k=2 tap falls off the cropped end
    x = torch.zeros(1, 4, 4, 1, dtype=F64)
    x[0, 3, 3, 0] = 1.0
    y = O.conv2d_transpose_same(x, w, None, 2)[0, :, :, 0]
    assert torch.equal(y[6:8, 6:8], w[:2, :2, 0, 0])


def test_transposed_conv_is_conv_input_gradient():
    x = torch.randn(2, 6, 6, 4, dtype=F64)
    w = torch.randn(3, 3, 5, 4, dtype=F64)
    y = O.conv2d_transpose_same(x, w, None, 2)
    xi = torch.randn(2, 12, 12, 5, dtype=F64, requires_grad=True)
    g, = torch.autograd.grad(O.conv2d_same(xi, w, None, 2), xi, x)
    assert (g - y).abs().max() < 1e-12
    w2 = torch.randn(2, 2, 5, 4, dtype=F64)
    g, = torch.autograd.grad(O.conv2d_same(xi, w2, None, 2), xi, x)
    assert (g - O.conv2d_transpose_same(x, w2, None, 2)).abs().max() < 1e-12


def test_zero_mask_equals_no_mask_graph():
    # SURVEY Q1 / KAT c-2: with mask == 0 and zero attention biases the live graph equals the as-written one
    p = O.init_params(O.generator_param_specs(64, True), 5, F64)
    x = torch.rand(1, 32, 32, 10, dtype=F64)
    y0 = O.generator_forward(p, x, None)
    y1 = O.generator_forward(p, x, torch.zeros(1, 32, 32, 1, dtype=F64))
    assert torch.equal(y0, y1)


def test_instance_norm_and_standardization():
    x = torch.randn(2, 8, 8, 3, dtype=F64) * 3 + 1
    y = O.instance_norm(x, torch.ones(3, dtype=F64), torch.zeros(3, dtype=F64))
    assert y.mean(dim=(1, 2)).abs().max() < 1e-12
    assert (y.var(dim=(1, 2), unbiased=False) - 1).abs().max() < 1e-5
    s, scale = O.per_image_standardization(x)
    assert torch.allclose(s * scale, x)
    assert torch.allclose(scale.flatten(), x.reshape(2, -1).std(dim=1, unbiased=False))
    z, zs = O.per_image_standardization(torch.zeros(1, 4, 4, 3, dtype=F64))
    assert zs.item() == pytest.approx(1 / 256.0)


def test_yuv_roundtrip_and_ssim_identity():
    x = torch.rand(1, 16, 16, 3, dtype=F64)
    assert (O.yuv_to_rgb(O.rgb_to_yuv(x)) - x).abs().max() < 1e-6
    assert O.ssim(x, x, 5.0).item() == pytest.approx(1.0)
    k = O.shmgan_oracle._gauss_kernel()
    assert k.sum().item() == pytest.approx(1.0) and tuple(k.shape) == (11, 11)


def test_keras_adam_first_step():
    # first step: m = (1-b1) g, v = (1-b2) g^2, lr_t = lr sqrt(1-b2)/(1-b1) => delta = lr * g/(|g| + eps*sqrt(1-b2)) ~ lr*sign(g)
    p = {"w": torch.tensor([1.0, -2.0], dtype=F64)}
    g = {"w": torch.tensor([0.5, -0.25], dtype=F64)}
    m = {"w": torch.zeros(2, dtype=F64)}
    v = {"w": torch.zeros(2, dtype=F64)}
    p, m, v = O.keras_adam_update(p, g, m, v, 0)
    assert torch.allclose(p["w"], torch.tensor([1.0 - 2e-5, -2.0 + 2e-5], dtype=F64), atol=1e-10)
    assert O.keras_adam_lr(10000) == pytest.approx(2e-5 * 0.95)


def test_pseudo_diffuse():
    a = [torch.randint(0, 256, (1, 4, 4, 3)).float() for _ in range(4)]
    ed = O.pseudo_diffuse_min4(*a)
    assert torch.equal(ed, torch.stack(a).min(dim=0).values)
