"""Multi-GPU plumbing on CPU: image-level sharding + the result/timing exchange bench.py does, world_size 2, gloo."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from vp8fix import GOLDEN, ROOT, sha


def test_shard_contiguous_partitions():
    from webp_decoder_b200.shard import shard_contiguous
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [list(shard_contiguous(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_shard_by_cost_balances_mixed_sizes():
    from webp_decoder_b200.shard import shard_by_cost
    rng = np.random.default_rng(0)
    costs = [int(c) for c in rng.choice([256, 1024, 3600, 8160, 14400, 32400], 200)]
    parts = shard_by_cost(costs, 8)
    assert sorted(sum(parts, [])) == list(range(200))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert shard_by_cost(costs, 8) == parts  # deterministic


def _rank_main(rank, world, port, names, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vp8fix import Oracle
    from webp_decoder_b200 import parse as P
    from webp_decoder_b200.shard import shard_by_cost
    from test_oracle import ParsedAsFrame
    pf = P.parse_batch([(GOLDEN / "webp" / n).read_bytes() for n in names], threads=2)
    costs = [pf.frames[i].mb_total for i in range(len(names))]
    mine = shard_by_cost(costs, world)[rank]
    orc = Oracle()  # stands in for the GPU on the CPU box: the exchange logic is what is under test
    local = {names[i]: sha(orc.decode_i420(ParsedAsFrame(pf.kfs[i], pf.frames[i]), True)) for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    t = torch.tensor([0.25 * (rank + 1)], dtype=torch.float64)  # per-rank elapsed; job time = max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        merged = {}
        for g in gathered:
            assert not (merged.keys() & g.keys())
            merged.update(g)
        q.put((merged, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharded_decode_gathers_every_frame_once(golden):
    names = [n for n in sorted(golden) if n.startswith("enc_")][:24]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = tmp.get_context("spawn")
    q = ctx.SimpleQueue()
    tmp.spawn(_rank_main, args=(2, port, names, q), nprocs=2, join=True)
    merged, tmax = q.get()
    assert tmax == 0.5
    assert merged == {n: golden[n]["yuvf"] for n in names}
