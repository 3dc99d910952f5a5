"""Single-layer convolution micro-benchmark (CUDA events, L2-cold by size): python tools/bench_conv.py N H W Cin Cout [k stride transposed] [--iters I]"""
import sys, os, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("dims", type=int, nargs="+")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--passes", default="fwd,dgrad,wgrad")
a = ap.parse_args()
N, H, W, Cin, Cout = a.dims[:5]
k = a.dims[5] if len(a.dims) > 5 else 3
s = a.dims[6] if len(a.dims) > 6 else 1
tr = bool(a.dims[7]) if len(a.dims) > 7 else False
c = ops.Conv("b", k, k, Cin, Cout, stride=s, transposed=tr, act=ops.ACT_LRELU, bias=True)
wshape = (k, k, Cout, Cin) if tr else (k, k, Cin, Cout)
c.w = torch.randn(wshape, device="cuda") * 0.05
c.b = torch.zeros(Cout, device="cuda")
c.dw = torch.zeros_like(c.w); c.db = torch.zeros(Cout, device="cuda")
x = torch.randn((N, H, W, Cin), device="cuda").bfloat16()
y = c.fwd(x, tc=True, version=1)
dy = torch.randn_like(y)
fl = c.flops(N, H, W)
nb = c.io_bytes(N, H, W, 2)
def timeit(fn):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters
for name, fn in (("fwd", lambda: c.fwd(x, y, tc=True, version=1)), ("dgrad", lambda: c.dgrad(dy, x.shape, None, tc=True, version=1)),
                 ("wgrad", lambda: c.wgrad(x, dy, tc=True))):
    if name not in a.passes.split(","): continue
    ms = timeit(fn)
    print("%s N=%d %dx%d %d->%d k%d s%d tr%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s algorithmic" % (name, N, H, W, Cin, Cout, k, s, tr, ms, fl / ms / 1e9, nb / ms / 1e6))
