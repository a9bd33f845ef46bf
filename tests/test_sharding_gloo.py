"""world_size-2 gloo test of the only multi-GPU logic on this path: utterance sharding (no data-path collective)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from zvxload import zvx
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = zvx.synth.batch_lengths(512, seed=13)
    mine = zvx.sharding.shard_utterances(lengths, world)[rank]
    # every rank derives the same assignment locally; the only exchange is bookkeeping for the test
    cnt = torch.tensor([len(mine), int(lengths[mine].sum())], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, cnt)
    mask = torch.zeros(512, dtype=torch.int64)
    mask[torch.from_numpy(mine)] = 1
    dist.all_reduce(mask)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((bool((mask == 1).all()), [g.tolist() for g in gathered], float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_partitions_and_balances():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29547, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert sum(g[0] for g in gathered) == 512
    loads = np.array([g[1] for g in gathered])
    assert abs(int(loads[0]) - int(loads[1])) <= 800
    assert tmax == 2.0
