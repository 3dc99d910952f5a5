"""One inference batch with nothing else (for ncu captures): python tools/prof_one.py [B S]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import model as M
B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (16, 512)
net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="bf16", allow_random_specseg=True).build()
img = torch.rand((B, S, S, 3), device="cuda")
net.inference_step(img)
torch.cuda.synchronize()
