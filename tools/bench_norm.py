"""Bandwidth micro-benchmark of the instance-norm kernels (CUDA events; tensors >> L2): python tools/bench_norm.py [N H W C]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import ops
dims = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else [80, 256, 256, 64]
N, H, W, C = dims
x = torch.randn((N, H, W, C), device="cuda").bfloat16()
dy = torch.randn((N, H, W, C), device="cuda").bfloat16()
dyp = torch.randn((N, H // 2, W // 2, C), device="cuda").bfloat16()
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
db = torch.zeros(C, device="cuda")
out = torch.empty_like(x); cat = torch.empty((N, H, W, 2 * C), device="cuda", dtype=torch.bfloat16)
sums = ops.inorm_stats(x)
e = x.numel() * 2
def timeit(name, fn, nbytes, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-34s %.3f ms  %6.0f GB/s (%.0f%% of 6544)" % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 65.44))
print("N=%d %dx%d C=%d bf16 (%.0f MB per tensor)" % (N, H, W, C, e / 1e6))
timeit("inorm_stats", lambda: ops.inorm_stats(x), e)
timeit("inorm_apply", lambda: ops.inorm_apply(x, sums, gamma, beta, out=out), 2 * e)
timeit("inorm_apply +pool -> cat slice", lambda: ops.inorm_apply(x, sums, gamma, beta, out=cat[..., C:], pooled=True), 2.25 * e)
timeit("inorm_bwd (stats+apply+dbias)", lambda: ops.inorm_bwd(x, sums, gamma, dy, None, dx=out, dbias=db), 5 * e)
timeit("inorm_bwd dyA+dyP", lambda: ops.inorm_bwd(x, sums, gamma, dy, dyp, dx=out, dbias=db), 5.5 * e)
timeit("act_bwd + dbias", lambda: ops.act_bwd(dy, x, 1, out=out, dbias=db), 3 * e)
