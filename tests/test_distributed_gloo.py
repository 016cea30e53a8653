"""N>1 host logic on CPU: world_size-2 gloo run of the game sharding, the
example gather to rank 0 and the statistics all-reduce."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
    from azb200.distributed import allreduce_game_stats, gather_examples_to_rank0, shard_games
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_games(8193, rank, world)
    n = 5 + 3 * rank                                   # ragged: ranks hold different numbers of examples
    obs = torch.full((n, 4, 6, 7), float(rank)) + torch.arange(n).view(n, 1, 1, 1)
    pi = torch.full((n, 7), float(rank))
    z = torch.full((n, 3), float(rank))
    g = gather_examples_to_rank0(obs, pi, z)
    st = allreduce_game_stats({"sims": 100 * (rank + 1), "games_played": 10 + rank, "peak_nodes": 7}, torch.device("cpu"))
    empty = gather_examples_to_rank0(torch.zeros(0, 4, 6, 7), torch.zeros(0, 7), torch.zeros(0, 3))
    if rank == 0:
        out.put(dict(shard=(first, count), obs=g[0].numpy(), pi=g[1].numpy(), z=g[2].numpy(), st=st,
                     empty=int(empty[0].shape[0])))
    else:
        assert g == (None, None, None)
        out.put(dict(shard=(first, count)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(r["shard"] for r in res)
    assert shards == [(0, 4097), (4097, 4096)]           # contiguous, balanced, covers all games
    r0 = next(r for r in res if "obs" in r)
    assert r0["obs"].shape == (5 + 8, 4, 6, 7)
    assert np.all(r0["pi"][:5] == 0) and np.all(r0["pi"][5:] == 1)   # rank order preserved
    assert np.array_equal(r0["obs"][5:, 0, 0, 0], 1 + np.arange(8))
    assert r0["st"]["sims"] == 300 and r0["st"]["games_played"] == 21 and r0["st"]["peak_nodes"] == 7
    assert r0["empty"] == 0
