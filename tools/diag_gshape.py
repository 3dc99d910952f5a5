import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from collections import OrderedDict
import torch
import oracle as O
from _util import F64, dev, rand, randn, rel_err, rms_err
from shmgan_b200 import nets

def run(fs, B, S, seed):
    p = O.init_params(O.generator_param_specs(fs, True), seed, F64, randomize_all=True)
    x, mask, dy = rand((B, S, S, 10), seed + 1), rand((2, S, S, 1), seed + 2), randn((B, S, S, 1), seed + 3)
    pr = OrderedDict((k, v.clone().requires_grad_(not k.endswith(("in_gamma", "in_beta")))) for k, v in p.items())
    reps = B // 2
    y = O.generator_forward(pr, x, mask.repeat(reps, 1, 1, 1))
    names = [k for k, v in pr.items() if v.requires_grad]
    grads = torch.autograd.grad((y * dy).sum(), [pr[k] for k in names])
    G = nets.Generator(fs, True, torch.float32, tensor_core=False)
    G.store.load(p)
    feats, saved = G.attention(dev(mask))
    yd, tape = G.forward(dev(x), feats, save=True)
    G.store.zero_grad()
    dattn = [torch.zeros_like(f) for f in feats]
    G.backward(tape, dev(dy), dattn, attn_nb=2)
    G.attention_backward(saved, dattn)
    got = G.store.export_grads()
    rows = sorted(((rel_err(got[k], w), k) for k, w in zip(names, grads) if float(w.abs().max()) > 0), reverse=True)
    print("fs %d B %d S %d seed %d fwd %.2e worst" % (fs, B, S, seed, rel_err(yd, y)), ["%s %.1e" % (k, e) for e, k in rows[:4]])

for seed in (1, 2, 3):
    run(8, 10, 64, seed)
run(8, 2, 64, 1)
run(16, 10, 32, 1)
