"""Multi-GPU host logic on CPU: world_size-2 gloo (no GPU needed).

The path shards by independent volume with no data-path collective (DESIGN.md, row e):
each rank owns a contiguous block of the batch (batch.shard_range), results stay
sharded, and the only collectives are the timing barrier and the max-over-ranks
reduction bench.py uses.  Here each rank deforms its shard with the CPU oracle
(standing in for the device), and rank 0 checks that the union equals the whole batch.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from elasticdeform_b200.batch import shard_range
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = 5
    rng = np.random.default_rng(42)                     # every rank draws the same batch
    Xs = [rng.random((12, 14, 16), dtype=np.float32) for _ in range(nb)]
    Ds = [rng.standard_normal((3, 3, 3, 3)) * 2 for _ in range(nb)]
    lo, hi = shard_range(nb, rank, world)
    mine = {b: O.deform_grid(Xs[b], Ds[b], order=1, impl="port") for b in range(lo, hi)}
    dist.barrier()
    t = torch.tensor([0.5 + rank], dtype=torch.float64)     # pretend per-rank step time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"range": (lo, hi), "keys": sorted(mine)})
    if rank == 0:
        full = [O.deform_grid(Xs[b], Ds[b], order=1, impl="port") for b in range(nb)]
        ok = all(np.array_equal(mine[b], full[b]) for b in mine)
        covered = sorted(k for g in gathered for k in g["keys"])
        q.put({"ok": ok, "covered": covered, "tmax": float(t.item()),
               "ranges": [g["range"] for g in gathered]})
    dist.destroy_process_group()


def test_shard_range_partitions():
    from elasticdeform_b200.batch import shard_range
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["ok"]
    assert res["covered"] == [0, 1, 2, 3, 4]
    assert res["ranges"] == [(0, 3), (3, 5)]
    assert res["tmax"] == 1.5
