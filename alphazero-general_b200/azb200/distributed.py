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


def train_steps_sharded(net, optimizer, loader, steps, value_loss_weight, device, group=None, sync_bn_stats=True,
                        bn_eval=False):
    """Data-parallel form of the loop body of NNetWrapper.train (NNetWrapper.py:131-165) over all ranks.

    Rank 0 owns the sample window and draws the batches exactly as the single-process loop does (``loader`` is its
    WindowLoader; the other ranks pass None); every step it broadcasts the batch, rank r computes the forward /
    backward pass of its contiguous slice with the losses normalised by the FULL batch size, the gradients are summed
    over ranks and every rank takes the same optimizer step -- so the parameters stay identical on all ranks and the
    step equals the single-process one up to the summation order of the gradient (and, in train mode, BatchNorm's
    per-rank batch statistics; the running statistics are averaged over ranks after the last step when
    ``sync_bn_stats``; ``bn_eval`` freezes BatchNorm, which makes the step independent of the split).  -> (mean policy loss, mean value loss) over the steps, as NNetWrapper.train reports them.
    """
    from .samples import loss_pi, loss_v
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    head = torch.zeros(6, dtype=torch.int64, device=device)      # batch rows, obs C/H/W, actions, values
    params = [p for p in net.parameters() if p.requires_grad]
    net.train(not bn_eval)
    it = None
    acc = torch.zeros(2, device=device, dtype=torch.float64)       # sum of batch-weighted losses, kept on the device
    n = 0
    for step in range(int(_bcast_int(steps if rank == 0 else 0, device, group))):
        if rank == 0:
            try:
                batch = next(it) if it is not None else None
            except StopIteration:
                batch = None
            if batch is None:                                  # a new epoch of the loader (new permutation)
                it = iter(loader)
                batch = next(it)
            boards, pis, vs = (t.to(device).contiguous() for t in batch)
            head.copy_(torch.tensor([boards.shape[0], *boards.shape[1:], pis.shape[1], vs.shape[1]], dtype=torch.int64))
        dist.broadcast(head, 0, group=group)
        b, c, h, w, na, nv = head.tolist()
        if rank != 0:
            boards = torch.empty((b, c, h, w), device=device)
            pis = torch.empty((b, na), device=device)
            vs = torch.empty((b, nv), device=device)
        for t in (boards, pis, vs):
            dist.broadcast(t, 0, group=group)
        first, count = shard_games(b, rank, world)             # same contiguous, balanced split as the games
        optimizer.zero_grad()
        stat = torch.zeros(2, device=device)
        if count:
            out_pi, out_v = net(boards[first:first + count])
            l_pi = loss_pi(pis[first:first + count], out_pi) * (count / b)      # sum over my rows / full batch size
            l_v = loss_v(vs[first:first + count], out_v, value_loss_weight) * (count / b)
            (l_pi + l_v).backward()
            stat[0], stat[1] = l_pi.detach(), l_v.detach()
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params] + [stat])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)               # one bucket: the net is < 25 MB
        off = 0
        for p in params:
            k = p.numel()
            p.grad = flat[off:off + k].view_as(p).clone()
            off += k
        optimizer.step()
        acc += flat[off:off + 2].double() * b
        n += b
    if sync_bn_stats and world > 1:
        for name, buf in net.named_buffers():
            if buf.dtype.is_floating_point:                    # running_mean / running_var
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
                buf /= world
    net.eval()
    return (float(acc[0]) / n, float(acc[1]) / n) if n else (0.0, 0.0)


def train_steps_local_windows(net, optimizer, window, iteration, args, value_loss_weight, device, group=None,
                              sync_bn_stats=True, bn_eval=False):
    """Data-parallel training without moving the examples (DDP NNetWrapper.train): every rank keeps the
    examples of its OWN games in its own SampleWindow (nothing is gathered to rank 0), draws train_batch_size / world
    rows per step from it and the gradients are averaged over the ranks -- the standard data-parallel estimator of the
    same loss (NNetWrapper.py:234-238) over the union of the shards.  The step count follows Coach.train's rule
    (Coach.py:452-478) on the GLOBAL sample counts.  Parameters stay identical on all ranks.
    -> (mean policy loss, mean value loss, steps, global samples in the window)."""
    from .samples import WindowLoader, loss_pi, loss_v
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    g = lambda k, d: (args[k] if k in args else d)
    its = window.window(iteration, args)
    sizes = torch.tensor([int(window.iters[i][0].shape[0]) for i in its] + [len(its)], device=device, dtype=torch.int64)
    n_its = torch.tensor([len(its)], device=device, dtype=torch.int64)
    dist.all_reduce(n_its, op=dist.ReduceOp.MIN, group=group)
    if int(n_its.item()) != len(its):
        raise RuntimeError("ranks disagree on the iterations in their sample windows")
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
    gsz = sizes[:-1].tolist()
    bs = int(g("train_batch_size", 1024))
    if not g("autoTrainSteps", True):
        steps = int(g("train_steps_per_iteration", 64))
    else:
        steps = ((sum(gsz) // len(gsz)) if g("averageTrainSteps", False) else gsz[-1]) // bs if gsz else 0
    local_bs = shard_games(bs, rank, world)[1]
    params = [p for p in net.parameters() if p.requires_grad]
    net.train(not bn_eval)
    loader = WindowLoader(window.tensors(its), max(local_bs, 1)) if its else None
    it = None
    acc = torch.zeros(3, device=device, dtype=torch.float64)
    for _ in range(steps):
        try:
            batch = next(it) if it is not None else None
        except StopIteration:
            batch = None
        if batch is None:
            it = iter(loader)
            batch = next(it)
        boards, pis, vs = batch
        optimizer.zero_grad()
        out_pi, out_v = net(boards)
        l_pi, l_v = loss_pi(pis, out_pi), loss_v(vs, out_v, value_loss_weight)
        (l_pi + l_v).backward()
        rows = float(boards.shape[0])
        stat = torch.stack([l_pi.detach() * rows, l_v.detach() * rows, torch.tensor(rows, device=device)])
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params] + [stat])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)               # one bucket: the net is < 25 MB
        off = 0
        for p in params:
            k = p.numel()
            p.grad = (flat[off:off + k] / world).view_as(p).clone()
            off += k
        optimizer.step()
        acc += flat[off:off + 3].double()
    if sync_bn_stats and world > 1:
        for name, buf in net.named_buffers():
            if buf.dtype.is_floating_point:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
                buf /= world
    net.eval()
    n = float(acc[2])
    return ((float(acc[0]) / n, float(acc[1]) / n) if n else (0.0, 0.0)) + (steps, sum(gsz))


def _bcast_int(v, device, group=None):
    t = torch.tensor([int(v)], dtype=torch.int64, device=device)
    dist.broadcast(t, 0, group=group)
    return int(t.item())
