"""Shared helpers of the GPU parity tests: seeded tensors, error norms, oracle <-> device conversion."""
import torch

import oracle as O

F64 = torch.float64


def gen(seed):
    return torch.Generator().manual_seed(seed)


def randn(shape, seed, scale=1.0):
    return torch.randn(shape, generator=gen(seed), dtype=F64) * scale


def rand(shape, seed):
    return torch.rand(shape, generator=gen(seed), dtype=F64)


def max_err(got, want):
    """max |got - want| / max |want| (both brought to fp64 on the host)."""
    g = got.detach().to("cpu", F64)
    w = want.detach().to("cpu", F64)
    assert g.shape == w.shape, (g.shape, w.shape)
    return float((g - w).abs().max() / (w.abs().max() + 1e-30))


def rms_err(got, want):
    """rms(got - want) / rms(want)."""
    g = got.detach().to("cpu", F64)
    w = want.detach().to("cpu", F64)
    assert g.shape == w.shape, (g.shape, w.shape)
    return float(((g - w) ** 2).mean().sqrt() / ((w ** 2).mean().sqrt() + 1e-30))


def rel_err(got, want):
    """The relative error every parity assertion uses: the LARGER of the max-norm reading (max |d| / max |want|) and the rms reading
    (rms d / rms want), so that a tolerance holds under both (max-norm alone is the most forgiving reading)."""
    return max(max_err(got, want), rms_err(got, want))


def cos_sim(got, want):
    g, w = got.detach().to("cpu", F64).reshape(-1), want.detach().to("cpu", F64).reshape(-1)
    return float((g @ w) / (g.norm() * w.norm() + 1e-300))


BF16_FACTOR = 1.5        # device error (vs the plain oracle) allowed as a multiple of what bf16 storage costs the oracle itself
BF16_FLOOR = 2e-2        # north_star's bf16 tolerance: tensors that bf16 storage barely moves are held to it directly
BF16_REPORT = {}         # test name -> rows (tensor, e_store, e_dev, cos_store, cos_dev), dumped to gpurun_out/ for DESIGN.md


def derived_bf16_grad_check(what, got: dict, plain: dict, stored: dict):
    """got / plain / stored: name -> gradient tensor from the device, the plain fp64 oracle and the fp64 oracle with bf16 storage points."""
    keys = [k for k, w in plain.items() if float(w.abs().max()) > 0]
    e_store = {k: rms_err(stored[k], plain[k]) for k in keys}
    med = sorted(e_store.values())[len(keys) // 2]
    rows, bad = [], []
    for k in keys:
        e_dev, c_dev, c_st = rms_err(got[k], plain[k]), cos_sim(got[k], plain[k]), cos_sim(stored[k], plain[k])
        rows.append((k, e_store[k], e_dev, c_st, c_dev))
        if e_dev > BF16_FACTOR * max(e_store[k], med) + BF16_FLOOR or (1 - c_dev) > BF16_FACTOR ** 2 * (1 - c_st) + 1e-3:
            bad.append(rows[-1])
    BF16_REPORT[what] = rows
    _dump_report()
    assert not bad, (what, "(tensor, e_store, e_dev, cos_store, cos_dev)", bad)
    return rows


def _dump_report():
    import json, os
    try:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "bf16_grad_bounds.json"), "w") as fh:
            json.dump({k: [{"tensor": r[0], "e_store": r[1], "e_dev": r[2], "cos_store": r[3], "cos_dev": r[4]} for r in v]
                       for k, v in BF16_REPORT.items()}, fh, indent=1)
    except OSError:
        pass


def dev(t, dtype=torch.float32):
    return t.to(dtype).cuda().contiguous()


def bf16_round(t):
    """fp64 tensor rounded to the bf16 grid (what the device sees in bf16 mode)."""
    return t.to(torch.float32).to(torch.bfloat16).to(F64)


def oracle_conv(x, w, b, k_stride, transposed, act):
    """Reference layer in the oracle: conv (+bias) (+activation)."""
    y = O.conv2d_transpose_same(x, w, b, k_stride) if transposed else O.conv2d_same(x, w, b, k_stride)
    if act == 1:
        y = O.leaky_relu(y)
    elif act == 2:
        y = torch.relu(y)
    elif act == 3:
        y = torch.sigmoid(y)
    return y
