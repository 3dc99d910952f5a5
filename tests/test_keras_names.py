"""The repo-name -> Keras-variable-name table (shmgan_b200/keras_names.py) against the only ground truth the reference ships:
the layer names, order and per-layer parameter counts of Generator_summary.txt / Discriminator_summary.txt / SpecSeg_summary.txt.
The numbers below are copied from those files (line numbers cited); /root/reference itself is not read."""
import math

from shmgan_b200 import keras_names as K
from shmgan_b200 import nets


def _layer_counts(specs, names):
    """Keras layer name -> number of parameters (kernel + bias / the four BN vectors); tfa IN variables are untracked."""
    out = {}
    for name, shape, kind in specs:
        if name.endswith(("in_gamma", "in_beta")):
            continue
        layer = names[name].split("/")[0]
        out[layer] = out.get(layer, 0) + math.prod(shape)
    return out


def test_generator_names_reproduce_the_summary():
    # Generator_summary.txt (as-written graph, 128 x 128): layer name, line, Param #
    want = [("conv2d", 7, 5824), ("conv2d_1", 39, 36928), ("conv2d_4", 73, 73856), ("conv2d_5", 105, 147584), ("conv2d_8", 139, 295168),
            ("conv2d_9", 171, 590080), ("conv2d_12", 205, 1180160), ("conv2d_13", 237, 2359808), ("conv2d_16", 271, 262656),
            ("conv2d_17", 303, 262656), ("conv2d_transpose", 335, 2359808), ("conv2d_18", 342, 4719104), ("conv2d_19", 374, 2359808),
            ("conv2d_transpose_1", 406, 1179904), ("conv2d_20", 413, 1179904), ("conv2d_21", 445, 590080),
            ("conv2d_transpose_2", 477, 295040), ("conv2d_22", 484, 295040), ("conv2d_23", 516, 147584),
            ("conv2d_transpose_3", 548, 73792), ("conv2d_24", 555, 73792), ("conv2d_25", 587, 36928), ("conv2d_26", 619, 65)]
    specs = nets.generator_specs(64, live_mask=False)
    names = K.generator_keras_names(64, live_mask=False)
    got = _layer_counts(specs, names)
    assert list(got) == [w[0] for w in want]                       # same layers in the same (creation) order
    for layer, _, n in want:
        assert got[layer] == n, layer
    assert sum(got.values()) == 18_525_569                          # Generator_summary.txt:621
    # the live-mask build re-uses the numbers the as-written graph skipped: conv2d_2,3 / 6,7 / 10,11 / 14,15
    live = K.generator_keras_names(64, live_mask=True)
    assert [live["attn%d%s.w" % (l, ab)] for l in (1, 2, 3, 4) for ab in "ab"] == \
        ["conv2d_%d/kernel:0" % i for i in (2, 3, 6, 7, 10, 11, 14, 15)]
    assert live["enc1a.w"] == "conv2d/kernel:0" and live["enc1a.b"] == "conv2d/bias:0" and live["out.w"] == "conv2d_26/kernel:0"


def test_discriminator_names_reproduce_the_summary():
    # Discriminator_summary.txt @128 x 128: conv2d_27 :9 1728, conv2d_28 :41 73728, conv2d_29 :73 294912, conv2d_30 :105 1179648,
    # conv2d_33 :139 4718592, conv2d_34 :175 9216, dense :177 81920; total :179 6,359,744
    want = [("conv2d_27", 1728), ("conv2d_28", 73728), ("conv2d_29", 294912), ("conv2d_30", 1179648), ("conv2d_33", 4718592),
            ("conv2d_34", 9216), ("dense", 81920)]
    got = _layer_counts(nets.discriminator_specs(128, 64, False), K.discriminator_keras_names(128, 64, False))
    assert list(got.items()) == want
    assert sum(got.values()) == 6_359_744
    live = K.discriminator_keras_names(128, 64, True)
    assert live["dattn_a.w"] == "conv2d_31/kernel:0" and live["dattn_b.b"] == "conv2d_32/bias:0"


def test_specseg_names_reproduce_the_summary():
    # SpecSeg_summary.txt: conv2d :8 160, conv2d_1 :14 2320, batch_normalization :17 64, conv2d_2 :22 4640, ..., conv2d_transpose :64 131200,
    # conv2d_10 :70 295040, ..., conv2d_18 :115 17; totals :118-120 1,942,801 / 1,941,809 trainable / 992 non-trainable
    names = K.specseg_keras_names()
    got = _layer_counts(nets.specseg_specs(), names)
    assert got["conv2d"] == 160 and got["conv2d_1"] == 2320 and got["batch_normalization"] == 64 and got["conv2d_2"] == 4640
    assert got["conv2d_transpose"] == 131200 and got["conv2d_10"] == 295040 and got["conv2d_18"] == 17
    assert sum(got.values()) == 1_942_801
    order = [n for n in dict.fromkeys(v.split("/")[0] for v in names.values())]
    assert order[:6] == ["conv2d", "conv2d_1", "batch_normalization", "conv2d_2", "conv2d_3", "batch_normalization_1"]
    assert order[15:19] == ["conv2d_transpose", "conv2d_10", "conv2d_11", "conv2d_transpose_1"]
    nontrainable = sum(math.prod(s) for n, s, _ in nets.specseg_specs() if n.endswith((".mean", ".var")))
    assert nontrainable == 992
    # get_weights() order: kernel, bias per conv; gamma, beta, moving_mean, moving_variance per BN
    assert list(names.values())[4:8] == ["batch_normalization/gamma:0", "batch_normalization/beta:0",
                                         "batch_normalization/moving_mean:0", "batch_normalization/moving_variance:0"]


def test_product_and_oracle_inventories_agree_but_are_checked_independently():
    """nets.*_specs (product) and oracle.*_param_specs (checker) are two statements of the Keras variable order; both are pinned to the summaries
    above / in test_oracle_kat.py, and they must agree with each other entry by entry."""
    import oracle as O
    for a, b in ((nets.generator_specs(64, True), O.generator_param_specs(64, True)),
                 (nets.discriminator_specs(256, 64, True), O.discriminator_param_specs(256, 64, True)),
                 (nets.specseg_specs(), O.specseg_param_specs())):
        assert [(n, tuple(s)) for n, s, _ in a] == [(n, tuple(s)) for n, s, _ in b]
