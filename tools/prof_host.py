"""Host-side profile of train_step (python/ctypes launch overhead): python tools/prof_host.py"""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import model as M, _lib
B, S = 16, 256
net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="bf16").build()
g = torch.Generator(device="cuda").manual_seed(1)
pol = [torch.rand((B, S, S, 3), generator=g, device="cuda") for _ in range(4)]
inp = pol + [net.calculate_estimate_diffuse(*pol)]
for _ in range(3):
    net.train_step(*inp)
torch.cuda.synchronize()
# host time of one step when the GPU queue is empty at entry (includes the final loss-table readback, which waits for the GPU)
for _ in range(2):
    t0 = time.perf_counter(); net.train_step(*inp); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("step: host %.1f ms, +sync %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
# host-only cost: skip the readback
orig = net.table.read
import types
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
net.train_step(*inp)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(25)
