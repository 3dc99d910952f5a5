"""Repo parameter name -> Keras variable name, for exchanging weights with a TensorFlow run of the reference.

Keras numbers layers of one type in CREATION order within a session: `conv2d`, `conv2d_1`, ... and `conv2d_transpose`, `conv2d_transpose_1`,
...  The reference builds G first, then D (ShmGANwithSSpecSeg.py:911-918), so D continues G's counters; SpecSeg is a separately saved model
with its own counters (SpecSeg_summary.txt).  The published summaries pin the order:

  Generator_summary.txt       conv2d :7, conv2d_1 :39, conv2d_4 :73, conv2d_5 :105, conv2d_8 :139, conv2d_9 :171, conv2d_12 :205, conv2d_13 :237,
                              conv2d_16 :271, conv2d_17 :303, conv2d_transpose :335, conv2d_18 :342, conv2d_19 :374, conv2d_transpose_1 :406, ...,
                              conv2d_26 :619 -- the gaps (conv2d_2,3,6,7,10,11,14,15) are the attention convs of attention_layer (:404-412), created
                              in that order but absent from the as-written graph; the live-mask build gives them those names back
  Discriminator_summary.txt   conv2d_27 :9 .. conv2d_30 :105, conv2d_33 :139 (31, 32 = the attention pair), conv2d_34 :175, dense :177
  SpecSeg_summary.txt         conv2d :8 .. conv2d_18 :115, batch_normalization :17 .. _4, conv2d_transpose :64 .. _3 :100

tfa InstanceNormalization variables are untracked in the reference; they are listed here under the names tfa gives them
(`instance_normalization[_k]/gamma:0`, `/beta:0`) so that a patched TF run can still be diffed.
"""
from __future__ import annotations

from collections import OrderedDict

from . import nets


def _suffix(base: str, i: int) -> str:
    return base if i == 0 else "%s_%d" % (base, i)


def _map(specs, first_conv=0, first_convT=0, first_in=0, first_bn=0, first_dense=0) -> "OrderedDict[str, str]":
    """Walks the specs in creation order; a layer = the run of entries sharing the name before the dot."""
    out: "OrderedDict[str, str]" = OrderedDict()
    counters = {"conv2d": first_conv, "conv2d_transpose": first_convT, "instance_normalization": first_in,
                "batch_normalization": first_bn, "dense": first_dense}
    layer_of = {}
    for name, shape, kind in specs:
        layer, var = name.split(".", 1)
        if var in ("in_gamma", "in_beta"):
            key = (layer, "in")
            if key not in layer_of:
                layer_of[key] = _suffix("instance_normalization", counters["instance_normalization"])
                counters["instance_normalization"] += 1
            out[name] = "%s/%s:0" % (layer_of[key], "gamma" if var == "in_gamma" else "beta")
            continue
        if layer.startswith("bn"):
            if layer not in layer_of:
                layer_of[layer] = _suffix("batch_normalization", counters["batch_normalization"])
                counters["batch_normalization"] += 1
            out[name] = "%s/%s:0" % (layer_of[layer], {"gamma": "gamma", "beta": "beta", "mean": "moving_mean", "var": "moving_variance"}[var])
            continue
        if layer not in layer_of:
            is_T = layer.startswith("up") or (layer[0] == "u" and layer[1:].isdigit())       # up1T..up4T (G), u6..u9 (SpecSeg)
            base = "dense" if layer == "dense" else ("conv2d_transpose" if is_T else "conv2d")
            layer_of[layer] = _suffix(base, counters[base])
            counters[base] += 1
        out[name] = "%s/%s:0" % (layer_of[layer], "kernel" if var == "w" else "bias")
    return out


def generator_keras_names(filter_size: int = 64, live_mask: bool = True) -> "OrderedDict[str, str]":
    """enc1a.w -> conv2d/kernel:0, enc1b.w -> conv2d_1/kernel:0, attn1a.w -> conv2d_2/kernel:0, ... out.w -> conv2d_26/kernel:0.
    With live_mask=False the attention names are skipped but their numbers stay consumed, exactly as in Generator_summary.txt."""
    full = _map(nets.generator_specs(filter_size, True))
    if live_mask:
        return full
    keep = {n for n, _, _ in nets.generator_specs(filter_size, False)}
    return OrderedDict((k, v) for k, v in full.items() if k in keep)


def discriminator_keras_names(image_size: int, filter_size: int = 64, live_mask: bool = True) -> "OrderedDict[str, str]":
    """d1.w -> conv2d_27/kernel:0 ... d4 -> conv2d_30, dattn_a/b -> conv2d_31/32, d5 -> conv2d_33, head -> conv2d_34, dense.w -> dense/kernel:0
    (D is built after G: 27 Conv2D layers and 18 instance norms already exist)."""
    full = _map(nets.discriminator_specs(image_size, filter_size, True), first_conv=27, first_in=18)
    if live_mask:
        return full
    keep = {n for n, _, _ in nets.discriminator_specs(image_size, filter_size, False)}
    return OrderedDict((k, v) for k, v in full.items() if k in keep)


def specseg_keras_names() -> "OrderedDict[str, str]":
    """c1a.w -> conv2d/kernel:0 ... u6.w -> conv2d_transpose/kernel:0 ... out.w -> conv2d_18/kernel:0; bn1.mean -> batch_normalization/moving_mean:0.
    The order of this dict is also the order of `keras_model.get_weights()`."""
    return _map(nets.specseg_specs())
