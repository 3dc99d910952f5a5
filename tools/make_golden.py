"""Generates tests/golden/*.npz: small seeded inputs and the CPU oracle's outputs for them (float64 arithmetic, stored as
float32/float64).  The reference itself (TensorFlow) cannot run in this image, so these are ORACLE-generated regression pins,
not reference outputs: they freeze the restated semantics (a change to the oracle that moves any number fails
tests/test_golden.py) and give the GPU tests a fixed target that does not depend on the oracle code at test time.

    python tools/make_golden.py          # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
F = torch.float64


def g(seed):
    return torch.Generator().manual_seed(seed)


def save(name, **arrs):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})


def convs():
    """Layer-level vectors: TF SAME padding for k3 s1, k3 s2 (pad 0,1), Conv2DTranspose k3 s2 / k2 s2 (crop at the end)."""
    out = {}
    for tag, (N, H, W, Ci, Co, k, s, tr) in {"k3s1": (2, 6, 5, 3, 4, 3, 1, False), "k3s2": (1, 8, 6, 2, 3, 3, 2, False),
                                              "k3s2_odd": (1, 7, 5, 2, 3, 3, 2, False), "t_k3s2": (1, 4, 3, 3, 2, 3, 2, True),
                                              "t_k2s2": (1, 3, 4, 2, 3, 2, 2, True), "k1s1": (1, 4, 4, 5, 2, 1, 1, False)}.items():
        x = torch.randn((N, H, W, Ci), generator=g(1), dtype=F)
        w = torch.randn((k, k, Co, Ci) if tr else (k, k, Ci, Co), generator=g(2), dtype=F) * 0.3
        b = torch.randn((Co,), generator=g(3), dtype=F) * 0.1
        xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
        y = O.conv2d_transpose_same(xr, wr, b, s) if tr else O.conv2d_same(xr, wr, b, s)
        dy = torch.randn(tuple(y.shape), generator=g(4), dtype=F)
        gx, gw = torch.autograd.grad((y * dy).sum(), [xr, wr])
        out.update({tag + "_x": x, tag + "_w": w, tag + "_b": b, tag + "_y": y, tag + "_dy": dy, tag + "_dx": gx, tag + "_dw": gw,
                    tag + "_cfg": np.array([N, H, W, Ci, Co, k, s, int(tr)])})
    save("conv_layers", **out)


def nets():
    fs, S = 4, 32
    Gp = O.init_params(O.generator_param_specs(fs, True), 1, F, randomize_all=True)
    Dp = O.init_params(O.discriminator_param_specs(S, fs, True), 2, F, randomize_all=True)
    Sp = O.init_params(O.specseg_param_specs(), 3, F, randomize_all=True)
    for k in Sp:
        if k.endswith(".var"):
            Sp[k] = Sp[k].abs() + 0.5
    x = torch.rand((2, S, S, 10), generator=g(5), dtype=F)
    mask = torch.rand((2, S, S, 1), generator=g(6), dtype=F)
    img = torch.rand((2, S, S, 3), generator=g(7), dtype=F)
    y1 = torch.rand((1, 16, 16, 1), generator=g(8), dtype=F)
    rf, cls = O.discriminator_forward(Dp, img, mask)
    inf = O.inference_step(Gp, Sp, img[:1, :16, :16].contiguous())
    save("nets", fs=fs, S=S, x=x, mask=mask, img=img, y1=y1,
         g_out=O.generator_forward(Gp, x, mask), g_out_nomask=O.generator_forward(Gp, x, None),
         d_rf=rf, d_cls=cls, specseg=O.specseg_forward(Sp, y1), inf_gen_rgb=inf["gen_rgb"], inf_mask=inf["mask"])


def step():
    fs, S, B = 4, 32, 1
    Gp = O.init_params(O.generator_param_specs(fs, True), 1, F, randomize_all=True)
    Dp = O.init_params(O.discriminator_param_specs(S, fs, True), 2, F, randomize_all=True)
    pol = [torch.rand((B, S, S, 3), generator=g(10 + i), dtype=F) for i in range(4)]
    origs = pol + [O.pseudo_diffuse_min4(*pol)]
    mask = torch.rand((B, S, S, 1), generator=g(20), dtype=F)
    bits = [True, False, True, False, False]
    L, gG, gD = O.train_step_grads(Gp, Dp, origs, mask, bits, 0.93, None, None, True, True, clip=False)
    names = ["total_Generator_loss", "total_Discriminator_loss", "total_Classification_loss", "G_gan_loss", "G_clsf_loss", "L1_loss_Gen",
             "ssim_cyc_loss", "Spec_loss", "content_loss", "style_loss", "total_NST_loss", "D4_RealFake_cyc", "D4_classification_loss"]
    arrs = {"loss_" + n: L[n] for n in names}
    arrs.update({"origs": torch.stack(origs), "mask": mask, "bits": np.array(bits), "T": 0.93, "gen_Y": L["gen_Y"], "gen_rgb": L["gen_rgb"]})
    arrs.update({"gG_" + k: v for k, v in gG.items() if k in ("enc1a.w", "dec4b.w", "up3T.w", "attn2b.w", "out.b")})
    arrs.update({"gD_" + k: v for k, v in gD.items() if k in ("d1.w", "d5.w", "dense.w", "dattn_b.w")})
    # Keras Adam: first update of one tensor from its raw gradient
    P, m, v = O.keras_adam_update({"w": Gp["dec4b.w"].clone()}, {"w": gG["dec4b.w"].clamp(-1, 1)}, {"w": torch.zeros_like(Gp["dec4b.w"])},
                                  {"w": torch.zeros_like(Gp["dec4b.w"])}, 0)
    arrs["adam_dec4b_w"] = P["w"]
    save("train_step", **arrs)


def prep():
    img = torch.rand((2, 8, 8, 3), generator=g(30), dtype=F)
    yuv, scale = O.per_image_standardization(O.rgb_to_yuv(img), True)
    pol = [(torch.rand((2, 5, 7, 3), generator=g(31 + i)) * 255).to(torch.uint8) for i in range(4)]
    a, b = torch.rand((2, 16, 16, 3), generator=g(40), dtype=F), torch.rand((2, 16, 16, 3), generator=g(41), dtype=F)
    save("prep", img=img, yuv=yuv, rgb_back=O.yuv_to_rgb(O.rgb_to_yuv(img)), pol=torch.stack(pol), ed=O.pseudo_diffuse_min4(*pol),
         ssim_a=a, ssim_b=b, ssim=O.ssim(O.rescale_01(a), O.rescale_01(b), 5.0), gram=O.gram_matrix(a))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    convs(); nets(); step(); prep()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
