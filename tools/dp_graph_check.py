"""Data-parallel train_step under CUDA-graph replay: torchrun --nproc-per-node 2 tools/dp_graph_check.py
Six steps eager and six steps with net.cuda_graph = True (same seeds): every rank's parameters must be bit-identical to rank 0's after each
run, and the graph run must track the eager run (losses to 1e-3, parameters to a few Adam steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from shmgan_b200 import model as M

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, B = 128, 4
g = torch.Generator(device="cuda").manual_seed(5 + rank)
pol = [torch.rand((B, S, S, 3), generator=g, device="cuda") for _ in range(4)]
patterns = [[True, False, True, False, False], [False, False, False, True, False]]


def run(graph):
    net = M.ShmGANwithSSpecSeg(M.default_args(image_size=S, batch_size=B, filter_size=32), dtype="bf16", allow_random_specseg=True).build()
    net.enable_data_parallel()
    net.cuda_graph = graph
    batch = pol + [net.calculate_estimate_diffuse(*pol)]
    losses = []
    for i in range(6):
        net.drop_bits, net.TARGET_LABELS = patterns[i % 2], 0.9 + 0.03 * i
        net.train_step(*batch)
        losses.append(net.total_Generator_loss)
    torch.cuda.synchronize()
    flats = [net.G.net.store.flat, net.D.net.store.flat]
    same = True
    for f in flats:
        ref = f.clone()
        dist.broadcast(ref, src=0)
        same = same and bool(torch.equal(ref, f))
    return net, losses, flats, same

e_net, e_loss, e_flat, e_same = run(False)
g_net, g_loss, g_flat, g_same = run(True)
dmax = max(float((a - b).abs().max()) for a, b in zip(e_flat, g_flat))
lerr = max(abs(a - b) / abs(a) for a, b in zip(e_loss, g_loss))
print("rank %d: eager ranks identical %s | graph ranks identical %s | graphs %d (cuda_graph still %s) | max |param eager - graph| %.2e | max loss rel diff %.2e"
      % (rank, e_same, g_same, len(g_net._graphs), g_net.cuda_graph, dmax, lerr), flush=True)
ok = e_same and g_same and dmax < 4e-4 and lerr < 5e-2
g_net.release_graphs()
dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
