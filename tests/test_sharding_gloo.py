"""world_size-2 gloo test of the request-sharding host logic (the N>1 path of bench.py / generate_batch)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csm_mlx_b200.sharding import gather_ragged, reduce_max, shard_indices


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5
    mine = shard_indices(n, rank, world)
    local = [torch.full((3 + i, 4), float(i)) for i in mine]  # request i -> (3+i, 4) tensor of value i
    got = gather_ragged(local, n)
    t = reduce_max(1.0 + rank)
    if rank == 0:
        ok = got is not None and len(got) == n and all(g.shape == (3 + i, 4) and bool((g == i).all()) for i, g in enumerate(got))
        q.put((ok, t))
    else:
        q.put((got is None, t))
    dist.destroy_process_group()


def test_shard_and_gather_two_ranks():
    assert shard_indices(5, 0, 2) == [0, 2, 4] and shard_indices(5, 1, 2) == [1, 3]
    assert sorted(shard_indices(64, 3, 8)) == list(range(3, 64, 8))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for ok, _ in res) and all(t == 2.0 for _, t in res)


def test_single_process_passthrough():
    out = gather_ragged([torch.ones(2), torch.zeros(3)], 2)
    assert len(out) == 2 and reduce_max(3.5) == 3.5
