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


def rel_err(got, want):
    """max |got - want| / max |want| (both brought to fp64 on the host)."""
    g = got.detach().to("cpu", F64)
    w = want.detach().to("cpu", F64)
    assert g.shape == w.shape, (g.shape, w.shape)
    return float((g - w).abs().max() / (w.abs().max() + 1e-30))


def rms_err(got, want):
    g = got.detach().to("cpu", F64)
    w = want.detach().to("cpu", F64)
    return float(((g - w) ** 2).mean().sqrt() / ((w ** 2).mean().sqrt() + 1e-30))


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
