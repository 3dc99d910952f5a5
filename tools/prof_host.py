"""Host-side cost of train_step (python / ctypes launch overhead) against its GPU time: python tools/prof_host.py
The loss-table readback (the only host<->device sync of a step) is stubbed out for the host-only measurement."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import model as M, _lib
B, S = 16, 256
net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="bf16", allow_random_specseg=True).build()
g = torch.Generator(device="cuda").manual_seed(1)
pol = [torch.rand((B, S, S, 3), generator=g, device="cuda") for _ in range(4)]
inp = pol + [net.calculate_estimate_diffuse(*pol)]
for _ in range(3):
    net.train_step(*inp)
torch.cuda.synchronize()
vals = net.table.read()
net.table.read = lambda: vals                     # no readback: train_step returns as soon as its launches are queued
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); net.train_step(*inp); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("step: host-only %.1f ms, GPU finished %.1f ms after the call started (%d launches)" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, 0))
l0 = _lib.launches()
pr = cProfile.Profile(); pr.enable()
net.train_step(*inp)
pr.disable()
torch.cuda.synchronize()
print("C-ABI calls per step:", _lib.launches() - l0)
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
