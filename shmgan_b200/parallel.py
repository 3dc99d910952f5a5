"""Batch data-parallel plumbing: one process per GPU, gradients of the flat per-network buffers summed with bucketed
all-reduces (NCCL over NVLink / NVSwitch) that run on the communicator's own stream while the rest of the backward keeps
computing.  The reference has no distributed code at all; the contract here is the one its
optimiser step implies: clip_by_value and Adam (ShmGANwithSSpecSeg.py:860-871) run on the AVERAGED gradient.

Works on CPU tensors with the gloo backend as well (that is how tests/ cover it without a GPU)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int):
    """Contiguous, even split of the global batch by sample (the path shards by independent polarimetric samples)."""
    if global_batch % world != 0:
        raise ValueError("global batch %d does not divide over %d ranks" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def bucket_ranges(n: int, bucket_elems: int, lo: int = 0):
    """[(a, b)] covering [lo, n) in buckets of at most bucket_elems elements, 4-element aligned starts."""
    bucket_elems = max(4, bucket_elems // 4 * 4)
    out, a = [], lo
    while a < n:
        b = min(n, a + bucket_elems)
        out.append((a, b))
        a = b
    return out


class GradReducer:
    def __init__(self, process_group=None, bucket_mb: float = 25.0):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.pending: List = []
        self.bytes_reduced = 0
        import os
        self.skip = os.environ.get("SHM_DP_NOREDUCE", "0") == "1"

    def broadcast_params(self, flats: Sequence[torch.Tensor], src: int = 0):
        for f in flats:
            dist.broadcast(f, src=src, group=self.pg)

    def reduce_async(self, flat: torch.Tensor, lo: int = 0, hi: int = -1):
        """Starts summing flat[lo:hi] over the ranks, one collective per bucket.  On CUDA the collectives are ordered after
        the work already queued on the current stream and run on NCCL's stream; `wait()` joins them."""
        hi = flat.numel() if hi < 0 else hi
        if self.skip:                                   # SHM_DP_NOREDUCE=1: diagnostic only (isolates straggler skew from communication)
            return
        for a, b in bucket_ranges(hi, self.bucket_elems, lo):
            self.pending.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
            self.bytes_reduced += (b - a) * flat.element_size()

    def wait(self):
        """Makes the current stream (CUDA) or the host (CPU/gloo) wait for every pending bucket."""
        for w in self.pending:
            w.wait()
        self.pending = []
