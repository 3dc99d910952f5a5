"""Does a tensor-bound wgrad kernel overlap with HBM-bound norm kernels on a second stream?  python tools/overlap_probe.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import ops

def mk(N, H, W, Cin, Cout):
    c = ops.Conv("b", 3, 3, Cin, Cout, act=ops.ACT_LRELU, bias=True)
    c.w = torch.randn((3, 3, Cin, Cout), device="cuda") * 0.05
    c.b = torch.zeros(Cout, device="cuda"); c.dw = torch.zeros_like(c.w); c.db = torch.zeros(Cout, device="cuda")
    x = torch.randn((N, H, W, Cin), device="cuda").bfloat16()
    dy = torch.randn((N, H, W, Cout), device="cuda").bfloat16()
    return c, x, dy

for shape in ((80, 64, 64, 512, 256), (80, 128, 128, 128, 128), (80, 256, 256, 64, 64)):
    c, x, dy = mk(*shape)
    N, H, W, C = 80, 256, 256, 64
    z = torch.randn((N, H, W, C), device="cuda").bfloat16(); g = torch.randn_like(z); out = torch.empty_like(z)
    gamma = torch.ones(C, device="cuda"); db = torch.zeros(C, device="cuda")
    sums = ops.inorm_stats(z)
    side = torch.cuda.Stream()
    def wg(): c.wgrad(x, dy, tc=True, bias_done=True)
    def nb(): ops.inorm_bwd(z, sums, gamma, g, None, dx=out, dbias=db)
    def timeit(fn, it=10):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it
    def both():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            wg()
        nb()
        torch.cuda.current_stream().wait_stream(side)
    tw, tn, tb = timeit(wg), timeit(nb), timeit(both)
    print("wgrad %s: %.3f ms | norm bwd 80x256x256x64: %.3f ms | both on two streams: %.3f ms (sum %.3f)" % (shape, tw, tn, tb, tw + tn))
