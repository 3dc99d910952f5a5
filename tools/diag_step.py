"""Diagnostic: fp32 train-step gradient errors (max-norm and L2) for one drop-bit pattern, device vs fp64 oracle, twice."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle as O
from _util import dev, rel_err, rms_err
import test_gpu_train_step as T

bits = [False] * 5
B, S, fs = 2, 64, 8
for rep in range(2):
    net, Gp, Dp, Sp, origs, noise, keep = T._setup("fp32", fs, B, S, bits)
    ds = [O.per_image_standardization(O.rgb_to_yuv(o), True)[0] for o in origs]
    mask = O.specseg_forward(Sp, ds[2][..., 0:1])
    L, gG, gD = O.train_step_grads(Gp, Dp, origs, mask, bits, 0.93, (noise[:B], noise[B:]), (keep[:B], keep[B:]), True, True, clip=False)
    net.train_step(*[dev(o) for o in origs])
    got = net.G.net.store.export_grads()
    rows = sorted(((rel_err(got[k], w), rms_err(got[k], w), k) for k, w in gG.items() if float(w.abs().max()) > 0), reverse=True)
    print("rep", rep)
    for em, e2, k in rows[:8]:
        print("   %-12s max %.3e l2 %.3e" % (k, em, e2))
    # where does the biggest up3T.b error sit?
    d = (got["up3T.b"].double() - gG["up3T.b"]).abs()
    print("   up3T.b err per channel", (d / gG["up3T.b"].abs().max()).tolist())
    for k in range(5):
        cy = torch.cat([L["cyc_Y"][k], L["avgCbCr"]], dim=3)
        flat = cy.reshape(B, -1)
        top = flat.topk(2, dim=1).values
        bot = (-flat).topk(2, dim=1).values
        print("   cyc%d max gap %s min gap %s" % (k, (top[:, 0] - top[:, 1]).tolist(), (bot[:, 0] - bot[:, 1]).tolist()))
