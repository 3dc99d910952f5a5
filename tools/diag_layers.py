"""Diagnostic: layer-by-layer bf16 forward error of the generator encoder vs the quantised oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from collections import OrderedDict
import torch
import oracle as O
from _util import F64, bf16_round, dev, rand, randn, rel_err, rms_err
from shmgan_b200 import nets

def run(dtype, tc, fs=64, B=8, S=64):
    q = O.bf16_storage if dtype == torch.bfloat16 else (lambda t: t)
    p = O.init_params(O.generator_param_specs(fs, True), 1, F64, randomize_all=True)
    p = OrderedDict((k, bf16_round(v)) for k, v in p.items())
    x, mask = bf16_round(rand((B, S, S, 10), 2)), bf16_round(rand((1, S, S, 1), 3))
    G = nets.Generator(fs, True, dtype, tensor_core=tc)
    G.store.load(p)
    feats, saved = G.attention(dev(mask, dtype))
    yd, tape = G.forward(dev(x, dtype), feats, save=True)
    attn = O.attention_features(p, mask.expand(B, S, S, 1), q=q)
    print("mode", dtype, "tc", tc)
    for l in range(4):
        print("  attn%d  l2 %.3e max %.3e" % (l + 1, rms_err(feats[l], attn[l][:1]), rel_err(feats[l], attn[l][:1])))
    h = x
    for lvl in range(1, 5):
        hin, za, sa, ya, zb, sb = tape["enc"][lvl - 1]
        oza = q(O.leaky_relu(O.conv2d_same(h, p[f"enc{lvl}a.w"], p[f"enc{lvl}a.b"])))
        oya = q(O.instance_norm(oza, p[f"enc{lvl}a.in_gamma"], p[f"enc{lvl}a.in_beta"]))
        ozb = q(O.leaky_relu(O.conv2d_same(oya, p[f"enc{lvl}b.w"], p[f"enc{lvl}b.b"])))
        y = O.instance_norm(ozb, p[f"enc{lvl}b.in_gamma"], p[f"enc{lvl}b.in_beta"])
        print("  enc%d  za l2 %.3e  ya l2 %.3e  zb l2 %.3e | fed-forward exact-input check:" % (lvl, rms_err(za, oza), rms_err(ya, oya), rms_err(zb, ozb)), end=" ")
        # single-op checks with the DEVICE's own input (isolates each kernel)
        za_h = za.double().cpu()
        oya2 = q(O.instance_norm(za_h, p[f"enc{lvl}a.in_gamma"], p[f"enc{lvl}a.in_beta"]))
        ozb2 = q(O.leaky_relu(O.conv2d_same(ya.double().cpu(), p[f"enc{lvl}b.w"], p[f"enc{lvl}b.b"])))
        print("IN %.3e conv %.3e" % (rms_err(ya, oya2), rms_err(zb, ozb2)))
        h = q(O.avg_pool2(y))

run(torch.bfloat16, True)
run(torch.bfloat16, False)
run(torch.float32, False)
