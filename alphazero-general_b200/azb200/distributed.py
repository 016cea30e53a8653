"""Multi-GPU plumbing: self-play shards by game (rank r owns games
[r*B, (r+1)*B) with its own node pool and RNG streams keyed by the global game
id), so the data path has no collective.  The only exchange is at the
iteration boundary: the variable-length (s, pi, z) examples of every rank are
gathered to rank 0 (Coach.saveIterationSamples, Coach.py:364-386, consumes
them there) and the scalar game statistics are all-reduced.  Works with the
nccl backend on device tensors and with gloo on host tensors."""
import torch
import torch.distributed as dist


def shard_games(total_games, rank, world):
    """[first, count) of the global games owned by ``rank`` (contiguous, balanced)."""
    base, extra = divmod(total_games, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_examples_to_rank0(obs, pi, z, group=None):
    """Gather per-rank example tensors [n_r, ...] to rank 0 in rank order.
    Returns (obs, pi, z) concatenated on rank 0, (None, None, None) elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = obs.device
    n = torch.tensor([obs.shape[0]], device=dev, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts) if counts else 0
    outs = []
    for t in (obs, pi, z):
        row = t.shape[1:]
        pad = torch.zeros((nmax,) + tuple(row), device=dev, dtype=t.dtype)
        pad[:t.shape[0]] = t
        if rank == 0:
            bufs = [torch.empty_like(pad) for _ in range(world)]
            dist.gather(pad, bufs, dst=0, group=group)
            outs.append(torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0))
        else:
            dist.gather(pad, None, dst=0, group=group)
            outs.append(None)
    return tuple(outs)


def allreduce_game_stats(stats, device, group=None):
    """Sum the engine counters (wins / draws / turns / sims / sum_depth ...) over ranks."""
    keys = sorted(k for k, v in stats.items() if isinstance(v, int))
    t = torch.tensor([stats[k] for k in keys], device=device, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = dict(stats)
    for k, v in zip(keys, t.tolist()):
        out[k] = int(v) if k not in ("peak_nodes",) else stats[k]
    return out
