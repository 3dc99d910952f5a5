"""Per-layer CUDA-event timing of one inference batch (SpecSeg mask + generator forward): python tools/prof_inference.py [B S]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import model as M, ops
B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (64, 512)
net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="bf16", allow_random_specseg=True).build()
img = torch.rand((B, S, S, 3), device="cuda")
for _ in range(2):
    net.inference_step(img)
ops.PROF = []
net.inference_step(img)
torch.cuda.synchronize()
rows = [(f, kind, name, fl, nb, e0.elapsed_time(e1)) for f, kind, name, fl, nb, e0, e1 in ops.PROF]
ops.PROF = None
tot = sum(r[5] for r in rows)
print("B=%d S=%d: %d profiled launches, %.2f ms" % (B, S, len(rows), tot))
for f, kind, name, fl, nb, t in sorted(rows, key=lambda r: -r[5])[:45]:
    print("%-34s %-8s %-10s %7.3f ms %5.1f%%  %7.0f TFLOP/s %6.0f GB/s" % (f, kind, name, t, 100 * t / tot, fl / t / 1e9, nb / t / 1e6))
