"""Data-parallel plumbing for N GPUs of one node: images shard contiguously across ranks, weights are
replicated, there is NO collective on the data path.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used
only for (a) the barrier / max-over-ranks timing bench.py needs and (b) the optional final gather of the top-k pairs
on rank 0 -- the north-star's "NCCL only for the optional final top-k gather"."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of a global batch for `rank`; the first (global_batch % world) ranks get one extra
    image, so ragged batches and world > batch (empty shards) are handled."""
    if world < 1 or not (0 <= rank < world) or global_batch < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Timing is reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_topk(idx: torch.Tensor, val: torch.Tensor, global_batch: int, dst: int = 0):
    """Gather per-rank top-k results ([b_r, k] int32 indices, [b_r, k] f32 probabilities; b_r may differ per rank) on
    rank `dst`, in global image order.  Returns (idx, val) on dst and (None, None) elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return idx, val
    world, rank = dist.get_world_size(), dist.get_rank()
    k = idx.shape[1]
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    cap = max(e - b for b, e in sizes)
    pad_i = torch.zeros(cap, k, dtype=torch.int32, device=idx.device)
    pad_v = torch.zeros(cap, k, dtype=torch.float32, device=val.device)
    pad_i[: idx.shape[0]] = idx
    pad_v[: val.shape[0]] = val
    out_i: Optional[List[torch.Tensor]] = [torch.empty_like(pad_i) for _ in range(world)] if rank == dst else None
    out_v: Optional[List[torch.Tensor]] = [torch.empty_like(pad_v) for _ in range(world)] if rank == dst else None
    dist.gather(pad_i, out_i, dst=dst)
    dist.gather(pad_v, out_v, dst=dst)
    if rank != dst:
        return None, None
    gi = torch.cat([out_i[r][: e - b] for r, (b, e) in enumerate(sizes)])
    gv = torch.cat([out_v[r][: e - b] for r, (b, e) in enumerate(sizes)])
    return gi, gv
