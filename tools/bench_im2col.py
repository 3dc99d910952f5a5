import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shmgan_b200 import ops
x = torch.rand((160, 256, 256, 3), device="cuda").bfloat16()
for _ in range(3): y = ops.im2col_k3s2(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y = ops.im2col_k3s2(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("im2col_k3s2 160x256x256x3 bf16: %.3f ms  %.0f GB/s" % (ms, (x.numel() * 2 + y.numel() * 2) / ms / 1e6))
g = torch.randn((96, 128, 128, 64), device="cuda").bfloat16()
for _ in range(3): d = ops.col2im_k3s2(g, 3, 256, 256, torch.bfloat16)
e0.record()
for _ in range(10): d = ops.col2im_k3s2(g, 3, 256, 256, torch.bfloat16)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("col2im_k3s2 96x128x128x64 -> 96x256x256x3 bf16: %.3f ms  %.0f GB/s (64 of the 128 bytes of a patch row are ever read)" % (ms, (g.numel() + d.numel() * 2) / ms / 1e6))
