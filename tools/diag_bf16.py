"""Diagnostic: per-tensor bf16-vs-oracle error of the generator / discriminator backward (max-norm and L2-norm relative),
against the plain fp64 oracle and against the oracle with the bf16 storage model (same quantisation points)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from collections import OrderedDict
import torch
import oracle as O
from _util import F64, bf16_round, dev, rand, randn, rel_err, rms_err
from shmgan_b200 import nets

def gen(dtype, fs, B, S, tc, q):
    p = O.init_params(O.generator_param_specs(fs, True), 1, F64, randomize_all=True)
    p = OrderedDict((k, bf16_round(v)) for k, v in p.items())
    x, mask, dy = bf16_round(rand((B, S, S, 10), 2)), bf16_round(rand((1, S, S, 1), 3)), bf16_round(randn((B, S, S, 1), 4))
    pr = OrderedDict((k, v.clone().requires_grad_(not k.endswith(("in_gamma", "in_beta")))) for k, v in p.items())
    xr = x.clone().requires_grad_()
    y = O.generator_forward(pr, xr, mask.expand(B, S, S, 1), q=q)
    names = [k for k, v in pr.items() if v.requires_grad]
    grads = torch.autograd.grad((y * dy).sum(), [pr[k] for k in names] + [xr])
    G = nets.Generator(fs, True, dtype, tensor_core=tc)
    G.store.load(p)
    feats, saved = G.attention(dev(mask, dtype))
    yd, tape = G.forward(dev(x, dtype), feats, save=True)
    G.store.zero_grad()
    dattn = [torch.zeros_like(f) for f in feats]
    dx = G.backward(tape, dev(dy, dtype), dattn, attn_nb=1, need_dx=True)
    G.attention_backward(saved, dattn)
    print("G %s tc=%s q=%s fwd max %.3e l2 %.3e | dx max %.3e l2 %.3e" % (dtype, tc, q is not None, rel_err(yd, y), rms_err(yd, y), rel_err(dx, grads[-1]), rms_err(dx, grads[-1])))
    got = G.store.export_grads()
    worst = sorted(((rms_err(got[k], w), rel_err(got[k], w), k) for k, w in zip(names, grads[:-1])), reverse=True)
    for e2, em, k in worst[:6]:
        print("   %-12s max %.3e l2 %.3e" % (k, em, e2))

def disc(dtype, fs, B, S, tc, q):
    p = O.init_params(O.discriminator_param_specs(S, fs, True), 5, F64, randomize_all=True)
    p = OrderedDict((k, bf16_round(v)) for k, v in p.items())
    x, mask, noise = bf16_round(rand((B, S, S, 3), 6)), bf16_round(rand((1, S, S, 1), 7)), bf16_round(randn((B, S, S, 3), 8) * 0.1)
    keep = (rand((B, S // 32, S // 32, fs * 16), 9) < 0.8).to(F64)
    d_rf, d_cls = randn((B, S // 32, S // 32, 1), 10), randn((B, 5), 11)
    pr = OrderedDict((k, v.clone().requires_grad_(not k.endswith(("in_gamma", "in_beta")))) for k, v in p.items())
    xr = x.clone().requires_grad_()
    rf, cls = O.discriminator_forward(pr, xr, mask.expand(B, S, S, 1), True, noise, keep, q=q)
    names = [k for k, v in pr.items() if v.requires_grad]
    grads = torch.autograd.grad((rf * d_rf).sum() + (cls * d_cls).sum(), [pr[k] for k in names] + [xr])
    D = nets.Discriminator(S, fs, True, dtype, tensor_core=tc)
    D.store.load(p)
    attn, saved = D.attention(dev(mask, dtype))
    rfd, clsd, tape = D.forward(dev(x, dtype), attn, dev(noise, dtype), dev(keep, dtype), save=True)
    D.store.zero_grad()
    dattn = torch.zeros_like(attn)
    dx = D.backward(tape, dev(d_rf), dev(d_cls), need_dx=True, dattn=dattn, attn_nb=1)
    D.attention_backward(saved, dattn)
    print("D %s tc=%s q=%s rf max %.3e l2 %.3e cls max %.3e | dx max %.3e l2 %.3e" % (dtype, tc, q is not None, rel_err(rfd, rf), rms_err(rfd, rf), rel_err(clsd, cls), rel_err(dx, grads[-1]), rms_err(dx, grads[-1])))
    got = D.store.export_grads()
    worst = sorted(((rms_err(got[k], w), rel_err(got[k], w), k) for k, w in zip(names, grads[:-1])), reverse=True)
    for e2, em, k in worst[:6]:
        print("   %-12s max %.3e l2 %.3e" % (k, em, e2))

bf = torch.bfloat16
gen(bf, 64, 8, 64, True, O.bf16_storage)
gen(bf, 64, 8, 64, True, None)
disc(bf, 64, 4, 128, True, O.bf16_storage)
disc(bf, 64, 4, 128, True, None)
