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


def _train_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
    from azb200 import nnet as aznet
    from azb200.distributed import train_steps_sharded
    from azb200.samples import WindowLoader
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = aznet.ResNet((4, 6, 7), 7, 3, num_channels=8, depth=1, value_dense_layers=(16,), policy_dense_layers=(16,))
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    loader = None
    if rank == 0:                                       # only rank 0 owns the sample window
        g = torch.Generator().manual_seed(5)
        n = 150                                         # 150 rows, batches of 64: a ragged last batch (22 rows)
        obs = torch.rand(n, 4, 6, 7, generator=g)
        pi = torch.softmax(torch.randn(n, 7, generator=g), 1)
        z = torch.softmax(torch.randn(n, 3, generator=g), 1)
        torch.manual_seed(9)
        loader = WindowLoader((obs, pi, z), 64)
    losses = train_steps_sharded(net, opt, loader, 5 if rank == 0 else 0, 1.5, torch.device("cpu"), bn_eval=True)
    out.put(dict(rank=rank, losses=losses, params=torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy(),
                 training=net.training))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_training_equals_single_process():
    """train_steps_sharded over two ranks = the loop body of NNetWrapper.train in one process on the same batches
    (BatchNorm frozen, so the only difference is the summation order of the gradient)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0]["params"], res[1]["params"])          # the ranks stay in lock-step
    assert res[0]["losses"] == res[1]["losses"] and not res[0]["training"]

    sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
    from azb200 import nnet as aznet
    from azb200.samples import WindowLoader, loss_pi, loss_v
    torch.manual_seed(0)
    net = aznet.ResNet((4, 6, 7), 7, 3, num_channels=8, depth=1, value_dense_layers=(16,), policy_dense_layers=(16,))
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(5)
    obs = torch.rand(150, 4, 6, 7, generator=g)
    pi = torch.softmax(torch.randn(150, 7, generator=g), 1)
    z = torch.softmax(torch.randn(150, 3, generator=g), 1)
    torch.manual_seed(9)
    loader = WindowLoader((obs, pi, z), 64)
    net.eval()                                                           # BatchNorm frozen as in the workers
    step, lp, lv, n = 0, 0.0, 0.0, 0
    while step < 5:
        for b, tp, tv in loader:
            if step == 5:
                break
            step += 1
            op, ov = net(b)
            l1, l2 = loss_pi(tp, op), loss_v(tv, ov, 1.5)
            opt.zero_grad(); (l1 + l2).backward(); opt.step()
            lp += float(l1.detach()) * len(b); lv += float(l2.detach()) * len(b); n += len(b)
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy()
    assert n == 64 + 64 + 22 + 64 + 64
    np.testing.assert_allclose(res[0]["params"], ref, rtol=0, atol=2e-6)
    np.testing.assert_allclose(res[0]["losses"], (lp / n, lv / n), rtol=1e-5)


def _local_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
    from azb200 import nnet as aznet
    from azb200.distributed import train_steps_local_windows
    from azb200.samples import SampleWindow
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = aznet.ResNet((4, 6, 7), 7, 3, num_channels=8, depth=1, value_dense_layers=(16,), policy_dense_layers=(16,))
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(5)                # every rank's shard holds the same rows in this test
    obs = torch.rand(150, 4, 6, 7, generator=g)
    pi = torch.softmax(torch.randn(150, 7, generator=g), 1)
    z = torch.softmax(torch.randn(150, 3, generator=g), 1)
    w = SampleWindow(device="cpu")
    w.add_iteration(1, obs, pi, z)
    torch.manual_seed(9)
    args = dict(train_batch_size=64, autoTrainSteps=False, train_steps_per_iteration=7)
    lp, lv, steps, gs = train_steps_local_windows(net, opt, w, 1, args, 1.5, torch.device("cpu"), bn_eval=True)
    out.put(dict(rank=rank, losses=(lp, lv), steps=steps, gs=gs,
                 params=torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy(), training=net.training))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_local_window_training():
    """train_steps_local_windows: every rank draws train_batch_size / world rows from its OWN window and the gradients
    are averaged.  With identical shards and RNG on both ranks the averaged gradient is the gradient of that half batch,
    so the run must equal one process training with batch size 32 on the same rows (BatchNorm frozen)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_local_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0]["params"], res[1]["params"]) and res[0]["losses"] == res[1]["losses"]
    assert res[0]["steps"] == 7 and res[0]["gs"] == 300 and not res[0]["training"]

    sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
    from azb200 import nnet as aznet
    from azb200.samples import WindowLoader, loss_pi, loss_v
    torch.manual_seed(0)
    net = aznet.ResNet((4, 6, 7), 7, 3, num_channels=8, depth=1, value_dense_layers=(16,), policy_dense_layers=(16,))
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(5)
    obs = torch.rand(150, 4, 6, 7, generator=g)
    pi = torch.softmax(torch.randn(150, 7, generator=g), 1)
    z = torch.softmax(torch.randn(150, 3, generator=g), 1)
    torch.manual_seed(9)
    loader = WindowLoader((obs, pi, z), 32)
    net.eval()
    step, lp, lv, n = 0, 0.0, 0.0, 0
    while step < 7:
        for b, tp, tv in loader:
            if step == 7:
                break
            step += 1
            op, ov = net(b)
            l1, l2 = loss_pi(tp, op), loss_v(tv, ov, 1.5)
            opt.zero_grad(); (l1 + l2).backward(); opt.step()
            lp += float(l1.detach()) * len(b); lv += float(l2.detach()) * len(b); n += len(b)
    ref = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy()
    np.testing.assert_allclose(res[0]["params"], ref, rtol=0, atol=2e-6)
    np.testing.assert_allclose(res[0]["losses"], (lp / n, lv / n), rtol=1e-5)
