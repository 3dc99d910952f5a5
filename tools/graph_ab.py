"""train_step eager vs CUDA-graph replay (net.cuda_graph): device ms per step and host ms per step.  python tools/graph_ab.py [B S steps]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shmgan_b200 import model as M

B, S, steps = ([int(v) for v in sys.argv[1:4]] + [16, 256, 10][len(sys.argv) - 1:])[:3]
g = torch.Generator(device="cuda").manual_seed(1)
pol = [torch.rand((B, S, S, 3), generator=g, device="cuda") for _ in range(4)]


def run(graph, bits):
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B), dtype="bf16", allow_random_specseg=True).build()
    net.cuda_graph, net.drop_bits = graph, bits
    batch = pol + [net.calculate_estimate_diffuse(*pol)]
    for _ in range(4):
        net.train_step(*batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        net.train_step(*batch)
    e1.record()
    host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    loss = net.total_Generator_loss
    del net
    torch.cuda.empty_cache()
    return e0.elapsed_time(e1) / steps, host, loss


for bits in ([True, False, True, False, False], [False] * 5):
    for graph in (False, True, False, True):
        ms, host, loss = run(graph, bits)
        print("bits %s  %-6s  %.2f ms/step on the device, %.2f ms/step of host time  (total_G %.4f)" % ("".join("1" if b else "0" for b in bits), "graph" if graph else "eager", ms, host, loss), flush=True)
