"""CPU tests of the N>1 host logic with the gloo backend, world_size 2 (and 3 for ragged shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, global_batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = pkg.dp()
    b, e = dp.shard_range(global_batch, rank, world)
    # fake per-rank results: index encodes the global image id so ordering can be verified
    idx = torch.arange(b, e, dtype=torch.int32).reshape(-1, 1).repeat(1, 5)
    val = idx.to(torch.float32) * 0.5
    gi, gv = dp.gather_topk(idx, val, global_batch, dst=0)
    t = dp.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((gi.tolist(), gv.tolist(), t))
    else:
        assert gi is None and gv is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,global_batch", [(2, 8), (2, 7), (3, 4), (2, 1)])
def test_shard_and_gather(world, global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    gi, gv, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [row[0] for row in gi] == list(range(global_batch))
    assert [row[0] for row in gv] == [0.5 * i for i in range(global_batch)]
    assert t == float(world)  # max over ranks of (1 + rank)


def test_shard_range_properties():
    dp = pkg.dp()
    for gb in (0, 1, 7, 256, 2048, 2049):
        for w in (1, 2, 3, 8):
            spans = [dp.shard_range(gb, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dp.shard_range(8, 2, 2)
