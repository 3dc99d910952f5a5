"""A/B of the big-halo convolution: single-CTA kernel (SHM_BIG2=0) vs the CTA-pair kernel, per layer shape, 80 images.
usage: [SHM_BIG2=0|1] [SHM_DEBUG=1] python tools/bench_big_pair.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shmgan_b200 import ops

def run(N, S, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(1)
    c = ops.Conv("t", 3, 3, cin, cout, act=1)
    c.w = torch.randn((3, 3, cin, cout), device="cuda", generator=g) * 0.03
    c.b = torch.zeros(cout, device="cuda")
    x = torch.randn((N, S, S, cin), device="cuda", generator=g).bfloat16()
    y = c.fwd(x, tc=True, version=1)
    for _ in range(3):
        c.fwd(x, y, tc=True, version=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        c.fwd(x, y, tc=True, version=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("big conv %4d -> %4d @%3dx%-3d N=%d: %.3f ms  %6.0f TFLOP/s   (SHM_BIG2=%s)" % (cin, cout, S, S, N, ms, c.flops(N, S, S) / ms / 1e9, os.environ.get("SHM_BIG2", "1")))

for shape in ((80, 128, 128, 128), (80, 128, 256, 128), (80, 64, 256, 256), (80, 64, 512, 256), (80, 32, 512, 512), (80, 32, 1024, 512)):
    run(*shape)
