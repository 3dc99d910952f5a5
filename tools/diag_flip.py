"""Diagnostic: is the up3T gradient outlier of the fp32 train-step test a LeakyReLU branch flip?  Re-runs the generator on the
step's own inputs and compares the up3T pre-activations (device fp32 vs fp64 on the device's own layer input)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle as O
from _util import dev
import test_gpu_train_step as T

bits = [False] * 5
B, S, fs = 2, 64, 8
net, Gp, Dp, Sp, origs, noise, keep = T._setup("fp32", fs, B, S, bits)
net.train_step(*[dev(o) for o in origs])
G = net.G.net
G.store.load(Gp)                      # weights before the update
mask = net.specular_candidate
feats, _ = G.attention(mask)
for name, x in (("G1", net.gen_input.contiguous()),):
    y, tape = G.forward(x, feats, save=True)
    for u in range(4):
        hin, cat = tape["dec"][u][0], tape["dec"][u][1]
        C = cat.shape[3] // 2
        pre = O.conv2d_transpose_same(hin.double().cpu(), Gp[f"up{u+1}T.w"], Gp[f"up{u+1}T.b"], 2)
        got = cat[..., :C].double().cpu()
        want = O.leaky_relu(pre)
        flips = ((got > 0) != (pre > 0))
        print(name, "up%dT" % (u + 1), "max|pre| %.3e" % float(pre.abs().max()), "min|pre| %.3e" % float(pre.abs().min()),
              "sign flips", int(flips.sum()), "at", flips.nonzero()[:4].tolist(), "pre there", pre[flips][:4].tolist(),
              "n(|pre|<1e-6)", int((pre.abs() < 1e-6).sum()))
