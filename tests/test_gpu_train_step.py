"""GPU parity: one full train_step (ShmGANwithSSpecSeg.py:467-875) and the inference body (test.py:218-250) through the
reference-shaped host class vs the CPU oracle: every published loss scalar, the raw gradients of both networks, and the
parameters after clip + Adam."""
from collections import OrderedDict

import pytest
import torch

import oracle as O
from _util import F64, bf16_round, derived_bf16_grad_check, dev, rand, randn, rel_err, rms_err

pytestmark = pytest.mark.gpu

SCALARS = ["total_Generator_loss", "total_Discriminator_loss", "total_Classification_loss", "G_gan_loss", "G_clsf_loss",
           "L1_loss_Gen", "ssim_cyc_loss", "Spec_loss", "content_loss", "style_loss", "total_NST_loss", "D4_RealFake_cyc",
           "D4_classification_loss"]


def _setup(dtype, fs, B, S, bits, live_mask=True):
    from shmgan_b200 import model as M
    args = M.default_args(image_size=S, batch_size=B, filter_size=fs)
    net = M.ShmGANwithSSpecSeg(args, dtype=dtype, live_mask=live_mask).build()
    Gp = O.init_params(O.generator_param_specs(fs, live_mask), 1, F64, randomize_all=True)
    Dp = O.init_params(O.discriminator_param_specs(S, fs, live_mask), 2, F64, randomize_all=True)
    Sp = O.init_params(O.specseg_param_specs(), 3, F64, randomize_all=True)
    for k in Sp:
        if k.endswith(".var"):
            Sp[k] = Sp[k].abs() + 0.5
    if dtype == "bf16":
        Gp, Dp, Sp = (OrderedDict((k, bf16_round(v)) for k, v in d.items()) for d in (Gp, Dp, Sp))
    net.G.net.store.load(Gp); net.D.net.store.load(Dp); net.SpecSeg.load(Sp)
    pol = [rand((B, S, S, 3), 10 + i) for i in range(4)]
    origs = pol + [O.pseudo_diffuse_min4(*pol)]
    noise = randn((2 * B, S, S, 3), 20) * 0.1
    keep = (rand((2 * B, S // 32, S // 32, fs * 16), 21) < 0.8).to(F64)
    if dtype == "bf16":
        noise = bf16_round(noise)
    net.drop_bits, net.TARGET_LABELS = bits, 0.93
    net.d_noise, net.d_keep = dev(noise), dev(keep)
    return net, Gp, Dp, Sp, origs, noise, keep


@pytest.mark.parametrize("bits", [[True, False, True, False, False], [False] * 5])
def test_train_step_fp32_parity(bits):
    B, S, fs = 2, 64, 8
    net, Gp, Dp, Sp, origs, noise, keep = _setup("fp32", fs, B, S, bits)
    ds = [O.per_image_standardization(O.rgb_to_yuv(o), True)[0] for o in origs]
    mask = O.specseg_forward(Sp, ds[2][..., 0:1])
    L, gG, gD = O.train_step_grads(Gp, Dp, origs, mask, bits, 0.93, (noise[:B], noise[B:]), (keep[:B], keep[B:]), True, True,
                                   clip=False)
    net.train_step(*[dev(o) for o in origs])
    assert rel_err(net.specular_candidate, mask) < 1e-3
    assert rel_err(net.gen_Y, L["gen_Y"]) < 1e-3 and rel_err(net.gen_rgb, L["gen_rgb"]) < 1e-3
    assert rel_err(net.cyc_genED_rgb, L["cyc_rgb"][4]) < 1e-3
    for name in SCALARS:
        assert getattr(net, name) == pytest.approx(float(L[name]), rel=1e-3, abs=1e-6), name
    gotD, gotG = net.D.net.store.export_grads(), net.G.net.store.export_grads()
    _check_grads(gotD, gD, _fp32_noise(Gp, Dp, origs, mask, bits, noise, keep, B, True)[1], "D grads")
    _check_grads(gotG, gG, _fp32_noise(Gp, Dp, origs, mask, bits, noise, keep, B, True)[0], "G grads")
    # parameters after clip_by_value(+-1) + Keras Adam (step 0), from the DEVICE's own raw gradients: isolates the update kernel
    for store, P, g in ((net.G.net.store, Gp, gotG), (net.D.net.store, Dp, gotD)):
        gc = OrderedDict((k, v.double().clamp(-1, 1)) for k, v in g.items())
        m = OrderedDict((k, torch.zeros_like(v)) for k, v in gc.items())
        v = OrderedDict((k, torch.zeros_like(vv)) for k, vv in gc.items())
        P2 = OrderedDict((k, P[k].float().double()) for k in gc)
        P2, m, v = O.keras_adam_update(P2, gc, m, v, 0, 2e-5, 0.5, 0.99, 1e-7)
        got = store.export()
        for k in P2:
            assert float((P2[k] - got[k].double()).abs().max()) < 2e-8, k      # the step itself is ~2e-5


_NOISE = {}


def _fp32_noise(Gp, Dp, origs, mask, bits, noise, keep, B, live):
    """Tolerance calibration = the conditioning of the problem itself, measured on the oracle: how far (per gradient tensor,
    max-norm relative) the float64 result moves when (a) the same oracle runs in float32 and (b) the float64 inputs are
    perturbed by 1e-7 relative (3 draws).  The random-init network amplifies rounding noise ~2x per conv block, the gradient
    crosses D, the cyclic G passes and G1, and a LeakyReLU pre-activation that sits within rounding of 0 (a few are expected
    per step) switches a local derivative 5x in ANY float32 evaluation, which shows up as ~1 % in the gradients upstream of it
    (tools/diag_step.py).  The 1e-3 fp32 target is therefore widened, per tensor, to 3x this measured sensitivity."""
    key = (tuple(bits), B, live)
    if key not in _NOISE:
        def run(c):
            _, g, d = O.train_step_grads(c(Gp), c(Dp), [c(o) for o in origs], c(mask), bits, 0.93, (c(noise[:B]), c(noise[B:])),
                                         (c(keep[:B]), c(keep[B:])), live, True, clip=False)
            return g, d
        g64, d64 = run(lambda t: t if torch.is_tensor(t) else OrderedDict((k, v) for k, v in t.items()))
        f = torch.float32
        runs = [run(lambda t: t.to(f) if torch.is_tensor(t) else OrderedDict((k, v.to(f)) for k, v in t.items()))]
        for draw in range(3):
            gen = torch.Generator().manual_seed(1000 + draw)
            pert = lambda t: t * (1 + 1e-7 * torch.randn(t.shape, generator=gen, dtype=t.dtype))
            _, g, d = O.train_step_grads(Gp, Dp, [pert(o) for o in origs], mask, bits, 0.93, (noise[:B], noise[B:]), (keep[:B], keep[B:]),
                                         live, True, clip=False)
            runs.append((g, d))
        nz = lambda ref: [k for k in ref if float(ref[k].abs().max()) > 0]
        _NOISE[key] = ({k: max(rel_err(r[0][k], g64[k]) for r in runs) for k in nz(g64)},
                       {k: max(rel_err(r[1][k], d64[k]) for r in runs) for k in nz(d64)})
    return _NOISE[key]


def _check_grads(got, want, noise32, what):
    """Every gradient tensor inside max(1e-3, 3 x measured sensitivity of the oracle) -- except that up to 10 % of the tensors
    may sit above it, below 2e-2: the signature of a LeakyReLU branch flip.  With ~2e6 pre-activations of scale ~0.5 and
    float32 rounding ~1e-7, about one pre-activation per step lands within rounding of 0; the device then takes the other
    branch than the float64 oracle at that one pixel and the gradients of that layer move by ~1 % (tools/diag_gshape.py: runs
    are either 4e-6 accurate everywhere or show one such outlier at a random layer)."""
    over, errs = {}, {}
    for k, w in want.items():
        if float(w.abs().max()) == 0:
            continue
        tol = min(2e-2, max(1e-3, 3.0 * noise32.get(k, 0.0)))
        errs[k] = rel_err(got[k], w)
        if errs[k] > tol:
            over[k] = (errs[k], tol)
    assert len(over) <= max(2, len(errs) // 10), (what, over)
    assert max(errs.values()) <= 2e-2, (what, {k: e for k, e in errs.items() if e > 2e-2})


def test_train_step_as_written_no_mask():
    """live_mask=False reproduces the reference as written (attention branch = exact zeros)."""
    B, S, fs, bits = 1, 64, 8, [False, True, False, False, True]
    net, Gp, Dp, Sp, origs, noise, keep = _setup("fp32", fs, B, S, bits, live_mask=False)
    ds = [O.per_image_standardization(O.rgb_to_yuv(o), True)[0] for o in origs]
    mask = O.specseg_forward(Sp, ds[2][..., 0:1])
    L, gG, gD = O.train_step_grads(Gp, Dp, origs, mask, bits, 0.93, (noise[:B], noise[B:]), (keep[:B], keep[B:]), False, True,
                                   clip=False)
    net.train_step(*[dev(o) for o in origs])
    for name in SCALARS:
        assert getattr(net, name) == pytest.approx(float(L[name]), rel=1e-3, abs=1e-6), name
    _check_grads(net.G.net.store.export_grads(), gG, _fp32_noise(Gp, Dp, origs, mask, bits, noise, keep, B, False)[0], "G grads")
    _check_grads(net.D.net.store.export_grads(), gD, _fp32_noise(Gp, Dp, origs, mask, bits, noise, keep, B, False)[1], "D grads")


def test_train_step_bf16_runs_and_tracks_oracle():
    """bf16 / tcgen05 mode: forward quantities within 2e-2 of the plain oracle, loss scalars within 5 %, and EVERY gradient tensor of both
    networks inside the bound derived in-test from what bf16 storage costs the oracle itself (tests/_util.py::derived_bf16_grad_check)."""
    B, S, fs, bits = 8, 64, 64, [True, False, True, False, False]
    net, Gp, Dp, Sp, origs, noise, keep = _setup("bf16", fs, B, S, bits)
    ds = [O.per_image_standardization(O.rgb_to_yuv(o), True)[0] for o in origs]
    mask = O.specseg_forward(Sp, bf16_round(ds[2][..., 0:1]))
    args = (Gp, Dp, origs, mask, bits, 0.93, (noise[:B], noise[B:]), (keep[:B], keep[B:]), True, True)
    L, gG, gD = O.train_step_grads(*args, clip=False)
    _, sG, sD = O.train_step_grads(*args, clip=False, q=O.bf16_storage_bwd)
    net.train_step(*[dev(o) for o in origs])
    assert rel_err(net.specular_candidate, mask) < 2e-2
    assert rel_err(net.gen_Y, L["gen_Y"]) < 2e-2 and rel_err(net.gen_rgb, L["gen_rgb"]) < 2e-2
    for name in SCALARS:
        assert getattr(net, name) == pytest.approx(float(L[name]), rel=5e-2, abs=1e-4), name
    assert torch.isfinite(net.G.net.store.flat).all() and torch.isfinite(net.D.net.store.flat).all()
    derived_bf16_grad_check("train_step G grads B=%d S=%d fs=%d" % (B, S, fs), net.G.net.store.export_grads(), gG, sG)
    derived_bf16_grad_check("train_step D grads B=%d S=%d fs=%d" % (B, S, fs), net.D.net.store.export_grads(), gD, sD)


def test_inference_step_matches_oracle():
    from shmgan_b200 import model as M
    B, S, fs = 2, 64, 16
    for dtype, tol in (("fp32", 1e-3), ("bf16", 2e-2)):
        net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype=dtype).build()
        Gp = O.init_params(O.generator_param_specs(fs, True), 1, F64, randomize_all=True)
        Sp = O.init_params(O.specseg_param_specs(), 3, F64, randomize_all=True)
        for k in Sp:
            if k.endswith(".var"):
                Sp[k] = Sp[k].abs() + 0.5
        net.G.net.store.load(Gp); net.SpecSeg.load(Sp)
        rgb = rand((B, S, S, 3), 30)
        want = O.inference_step(Gp, Sp, rgb)
        got = net.inference_step(dev(rgb))
        assert rel_err(net.specular_candidate, want["mask"]) < tol
        assert rel_err(got, want["gen_rgb"]) < tol
        agree = ((net.specular_candidate.cpu() > 0.5) == (want["mask"] > 0.5)).double().mean()
        near = ((want["mask"] - 0.5).abs() < (1e-4 if dtype == "fp32" else 5e-3)).double().mean()
        assert float(agree) >= 0.999 - float(near)


def test_inference_step_full_width_thin_first_layers():
    """filter_size 64 at 64 x 64: the inference body runs SpecSeg's 16/32-channel levels, the attention first convs (1 -> 64 / 128) and
    enc1a (10 -> 64) on the thin tensor-core kernels (16-channel pixel rows); the training-shaped path (64-channel padding) must agree."""
    from shmgan_b200 import model as M
    B, S, fs = 2, 64, 64
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype="bf16").build()
    Gp = O.init_params(O.generator_param_specs(fs, True), 1, F64, randomize_all=True)
    Sp = O.init_params(O.specseg_param_specs(), 3, F64, randomize_all=True)
    for k in Sp:
        if k.endswith(".var"):
            Sp[k] = Sp[k].abs() + 0.5
    Gp, Sp = (OrderedDict((k, bf16_round(v)) for k, v in d.items()) for d in (Gp, Sp))
    net.G.net.store.load(Gp); net.SpecSeg.load(Sp)
    assert net.G.net.in_channels(B, S, S, infer=True) == 16 and net.G.net.in_channels(B, S, S) == 16
    rgb = rand((B, S, S, 3), 31)
    want = O.inference_step(Gp, Sp, rgb)
    got = net.inference_step(dev(rgb)).clone()
    assert rel_err(net.specular_candidate, want["mask"]) < 2e-2
    assert rel_err(got, want["gen_rgb"]) < 2e-2
    thin = net.G.net.thin
    net.G.net.thin = {}                                  # same weights through the 64-channel padded first layers
    ref = net.inference_step(dev(rgb))
    net.G.net.thin = thin
    assert rel_err(got, ref.double().cpu()) < 1e-2


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_inference_step_cyclic_matches_restated_test_py(dtype, tol):
    """inference_step(cyclic=True) = test.py:252-297: five more generator passes whose non-target slots carry the R channel of gen_rgb
    (Q11), re-joined with the image's own CbCr; and gen_rgb_output (test.py:246-249) scales by the running mean of EVERY standardisation
    scale seen so far (self.stddev_arr is never cleared), checked over two consecutive calls."""
    from shmgan_b200 import model as M
    B, S, fs = 2, 64, 16
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype=dtype).build()
    Gp = O.init_params(O.generator_param_specs(fs, True), 1, F64, randomize_all=True)
    Sp = O.init_params(O.specseg_param_specs(), 3, F64, randomize_all=True)
    for k in Sp:
        if k.endswith(".var"):
            Sp[k] = Sp[k].abs() + 0.5
    if dtype == "bf16":
        Gp, Sp = (OrderedDict((k, bf16_round(v)) for k, v in d.items()) for d in (Gp, Sp))
    net.G.net.store.load(Gp); net.SpecSeg.load(Sp)
    history = []
    for call, seed in enumerate((32, 33)):
        rgb = rand((B, S, S, 3), seed) * (1.0 if call == 0 else 0.4)      # a darker second batch: its scales differ from the first
        want = O.inference_step(Gp, Sp, rgb, cyclic=True, stddev_history=history)
        history.append(want["scale"])
        got = net.inference_step(dev(rgb), cyclic=True)
        assert rel_err(got, want["gen_rgb"]) < tol
        names = ["cyc_gen0_rgb", "cyc_gen45_rgb", "cyc_gen90_rgb", "cyc_gen135_rgb", "cyc_genED_rgb"]
        for k in range(5):
            assert rel_err(getattr(net, names[k]), want["cyc_rgb"][k]) < tol, names[k]
            assert rel_err(net.cyc_rgb[k], want["cyc_rgb"][k]) < tol
        assert rel_err(net.gen_rgb_output, want["gen_rgb_output"]) < tol
        assert net.stddev_arr.mean() == pytest.approx(float(torch.cat([h.reshape(-1) for h in history]).mean()), rel=1e-5)


def test_train_step_publishes_gen_rgb_output_and_lazy_losses():
    """gen_rgb_output (:548-551) = yuv_to_rgb(gen_YCbCr * mean(stddev_arr) * 255) over the five per-image scales of every step so far; the loss
    scalars are read back on first access (no host synchronisation inside train_step) and a step's scalars replace the previous step's."""
    B, S, fs, bits = 2, 64, 8, [False, True, False, False, False]
    net, Gp, Dp, Sp, origs, noise, keep = _setup("fp32", fs, B, S, bits)
    scales = torch.stack([O.per_image_standardization(O.rgb_to_yuv(o), True)[1].reshape(-1) for o in origs])
    net.train_step(*[dev(o) for o in origs])
    assert net._pending_losses is not None and "total_Generator_loss" not in net.__dict__      # not read back yet
    want = net.gen_rgb.double().cpu() * float(scales.mean()) * 255.0
    assert rel_err(net.gen_rgb_output, want) < 1e-5
    first = net.total_Generator_loss
    assert net._pending_losses is None and first == net.__dict__["total_Generator_loss"]
    net.train_step(*[dev(o) for o in origs])
    assert "total_Generator_loss" not in net.__dict__ and net.total_Generator_loss != first    # weights moved: a new value
    assert len(net.stddev_arr) == 10 * B                                                       # 5 images x B samples x 2 steps


def test_random_specseg_weights_must_be_asked_for():
    """The reference cannot run without specsegv3_chkpt.h5 (:930-931): a live-mask step on SpecSeg's random initialisers is refused unless the
    caller opts in; loading weights (any of the importers) or a checkpoint that holds SpecSeg lifts the refusal."""
    from shmgan_b200 import model as M
    from shmgan_b200.keras_names import specseg_keras_names
    B, S, fs = 1, 64, 8
    rgb = dev(rand((B, S, S, 3), 40))
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype="fp32").build()
    with pytest.raises(RuntimeError, match="SpecSeg"):
        net.inference_step(rgb)
    with pytest.raises(RuntimeError, match="SpecSeg"):
        net.train_step(rgb, rgb, rgb, rgb, rgb)
    M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype="fp32", live_mask=False).build().inference_step(rgb)
    M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=fs), dtype="fp32", allow_random_specseg=True).build().inference_step(rgb)
    # Keras get_weights() order -> named parameters
    Sp = O.init_params(O.specseg_param_specs(), 3, F64, randomize_all=True)
    for k in Sp:
        if k.endswith(".var"):
            Sp[k] = Sp[k].abs() + 0.5
    assert list(Sp) == list(specseg_keras_names())
    net.SpecSeg.load_keras_weights([v.numpy() for v in Sp.values()])
    assert net.SpecSeg.loaded
    got = net.inference_step(rgb)
    Gp = {k: v.double() for k, v in net.G.net.store.export().items()}
    assert rel_err(got, O.inference_step(Gp, Sp, rgb.double().cpu())["gen_rgb"]) < 1e-3
    with pytest.raises(ValueError):
        net.SpecSeg.load_keras_weights([v.numpy() for v in Sp.values()][:-1])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_train_step_cuda_graph_replay_matches_eager(dtype):
    """net.cuda_graph = True: the step is captured once per drop-bit pattern and replayed.  Six steps with two alternating bit patterns, a fresh
    TARGET_LABELS draw every step (ShmGANwithSSpecSeg.py:986) and moving Adam / Philox counters must leave the same parameters, losses and
    published tensors as six eager steps (fp32 atomics make the two runs agree to rounding, not bitwise); the replayed steps must also count
    their kernels (gpu_launches evidence of bench.py)."""
    from shmgan_b200 import _lib, model as M
    S, B = 64, 2
    g = torch.Generator().manual_seed(77)
    pol = [torch.rand((B, S, S, 3), generator=g).cuda() for _ in range(4)]
    batch = pol + [torch.minimum(torch.minimum(pol[0], pol[1]), torch.minimum(pol[2], pol[3]))]
    patterns = [[True, False, True, False, False], [False, False, False, True, False]]
    Ts = [0.9, 1.1, 0.83, 1.17, 0.95, 1.02]

    def run(graph):
        net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=16), dtype=dtype, allow_random_specseg=True).build()
        net.noise_seed, net.cuda_graph = 11, graph                  # two eager warm-up steps, captures at steps 3 and 4, replays at 5 and 6
        p0 = {n: getattr(net, n).net.store.flat.clone() for n in ("G", "D")}
        losses, counts = [], []
        for i, T in enumerate(Ts):
            net.drop_bits, net.TARGET_LABELS = patterns[i % 2], T
            n0 = _lib.launches()
            net.train_step(*batch)
            counts.append(_lib.launches() - n0)
            losses.append((net.total_Generator_loss, net.D1_classification_loss, net.D2_RealFake_target))
        torch.cuda.synchronize()
        move = {n: (getattr(net, n).net.store.flat - p0[n]).double() for n in ("G", "D")}
        return net, losses, counts, move

    # Two eager runs differ too: fp32 atomics reorder the weight-gradient sums, and Adam's first steps move a parameter by ~lr * 5 * sign(g), so
    # a gradient that is pure rounding noise (a conv bias in front of an instance norm) takes a different +-2e-5 walk in every run.  The graph
    # run is held to 3 x what the second eager run measures (+ a floor), on the movement of the parameters over the six steps.
    eager, le, ce, me = run(False)
    eager2, le2, _, me2 = run(False)
    graph, lg, cg, mg = run(True)
    assert len(graph._graphs) == 2 and graph.step_count == eager.step_count == len(Ts)
    assert graph.G.net.store.step == eager.G.net.store.step and graph.D.net.store.step == eager.D.net.store.step
    assert min(cg[4:]) > 100 and cg[4] == cg[2] and cg[5] == cg[3]      # the replays (steps 5, 6) count the kernels recorded at capture
    tol = 2e-4 if dtype == "fp32" else 5e-2
    for a, b in zip(le, lg):
        for x, y in zip(a, b):
            assert y == pytest.approx(x, rel=tol, abs=tol * 1e-2)
    for name in ("G", "D"):
        base = float(me[name].norm())
        noise_l2, noise_max = float((me2[name] - me[name]).norm()) / base, float((me2[name] - me[name]).abs().max())
        got_l2, got_max = float((mg[name] - me[name]).norm()) / base, float((mg[name] - me[name]).abs().max())
        # L2 of the movement: systematic errors (a stale lr_t, TARGET_LABELS or Philox offset) show here; a handful of +-2e-5 sign flips do not.
        # Max-norm: whether ONE such flip happens is chance in any pair of runs, so it is held to the physical bound instead -- two walks of
        # six Adam steps of at most ~lr = 2e-5 each can end at most 2.4e-4 apart.
        assert base > 0 and got_l2 <= 3.0 * noise_l2 + 5e-3 and got_max <= 2.5e-4, (name, got_l2, noise_l2, got_max, noise_max)
    assert rel_err(graph.gen_rgb, eager.gen_rgb.double().cpu()) < (1e-3 if dtype == "fp32" else 5e-2)
    assert rel_err(graph.cyc_genED_rgb, eager.cyc_genED_rgb.double().cpu()) < (1e-3 if dtype == "fp32" else 5e-2)


def test_inference_step_cuda_graph_replay_matches_eager():
    """inference_step under net.cuda_graph: the first call is eager, the second records, later calls replay -- on NEW input images each time;
    after a training step the graph recorded against the old weights is dropped and a new one recorded."""
    from shmgan_b200 import _lib, model as M
    S, B = 64, 2
    g = torch.Generator().manual_seed(5)
    imgs = [torch.rand((B, S, S, 3), generator=g).cuda() for _ in range(5)]
    mk = lambda: M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=16), dtype="bf16", allow_random_specseg=True).build()
    eager, graph = mk(), mk()
    graph.cuda_graph = True
    for i, img in enumerate(imgs[:4]):
        want = eager.inference_step(img, cyclic=(i == 3)).clone()
        n0 = _lib.launches()
        got = graph.inference_step(img, cyclic=(i == 3))
        assert _lib.launches() - n0 > 50
        # same kernels in the same order; only the fp64 atomics of the instance-norm statistics may land in another order (a bf16 ulp here and there)
        assert rel_err(got, want.double().cpu()) < 2e-2, i
        assert rel_err(graph.specular_candidate, eager.specular_candidate.double().cpu()) < 2e-2
    assert abs(eager.stddev_arr.mean() - graph.stddev_arr.mean()) < 1e-12 and len(eager.stddev_arr) == len(graph.stddev_arr)
    assert sum(1 for k in graph._graphs if k[0] == "infer") == 1      # the cyclic call (first of its kind) ran eagerly
    pol = [torch.rand((B, S, S, 3), generator=g).cuda() for _ in range(4)]
    batch = pol + [torch.minimum(torch.minimum(pol[0], pol[1]), torch.minimum(pol[2], pol[3]))]
    for net in (eager, graph):
        net.cuda_graph = False
        net.drop_bits, net.noise_seed = [False] * 5, 3
        net.train_step(*batch)
    graph.cuda_graph = True
    for _ in range(3):                                          # eager (already warm) -> would replay a STALE graph if the key ignored the weights
        got = graph.inference_step(imgs[4])
    assert rel_err(got, eager.inference_step(imgs[4]).double().cpu()) < 5e-2
    assert sum(1 for k in graph._graphs if k[0] == "infer") == 1
