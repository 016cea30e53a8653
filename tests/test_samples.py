"""SampleWindow / WindowLoader (azb200/samples.py) against the input side of Coach.train (Coach.py:436-520):
the batches equal those of DataLoader(ConcatDataset(TensorDataset...), shuffle=True) under the same RNG state, the
window and step arithmetic equal the reference's, the three-file format round-trips."""
import numpy as np
import pytest
import torch
from torch.utils.data import ConcatDataset, DataLoader, TensorDataset

from azb200.samples import SampleWindow, WindowLoader, history_window, loss_pi, loss_v


def _iters(sizes, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {i + 1: (torch.rand(n, 4, 6, 7, generator=g), torch.rand(n, 7, generator=g), torch.rand(n, 3, generator=g))
            for i, n in enumerate(sizes)}


@pytest.mark.parametrize("bs", [64, 100])
def test_loader_yields_the_dataloader_batches(bs):
    its = _iters([130, 77, 201])
    w = SampleWindow(device="cpu")
    for i, t in its.items():
        w.add_iteration(i, *t)
    ref = DataLoader(ConcatDataset([TensorDataset(*its[i]) for i in sorted(its)]), batch_size=bs, shuffle=True, num_workers=0)
    mine = WindowLoader(w.tensors(sorted(its)), bs)
    assert len(mine) == len(ref)
    for seed in (0, 5):
        torch.manual_seed(seed)
        want = [b for _ in range(2) for b in ref]           # two epochs, as NNetWrapper.train loops over `batches`
        torch.manual_seed(seed)
        got = [b for _ in range(2) for b in mine]
        assert len(want) == len(got)
        for a, b in zip(want, got):
            for x, y in zip(a, b):
                assert torch.equal(x, y)


def test_window_and_step_arithmetic():
    args = dict(minTrainHistoryWindow=4, maxTrainHistoryWindow=20, trainHistoryIncrementIters=2, train_batch_size=32,
                autoTrainSteps=True, averageTrainSteps=False, train_steps_per_iteration=64)
    for it in range(1, 60):
        want = min(max(4, (it + 4) // 2), 20)                # Coach.py:510-516
        assert history_window(it, args) == want
    w = SampleWindow(device="cpu")
    sizes = [100, 330, 65, 40, 900, 31, 500]
    for i, t in _iters(sizes).items():
        w.add_iteration(i, *t)
    assert w.window(7, args) == [2, 3, 4, 5, 6, 7]           # range(max(1, 7 - 5), 8)
    assert w.window(3, args) == [1, 2, 3]
    assert w.train_steps([2, 3, 4, 5, 6, 7], args) == 500 // 32                     # the newest iteration's samples
    assert w.train_steps([2, 3, 4], dict(args, averageTrainSteps=True)) == ((330 + 65 + 40) // 3) // 32
    assert w.train_steps([1, 2], dict(args, autoTrainSteps=False)) == 64
    assert w.train_steps([1, 2], args, train_on_all=True) == 430 // 32
    w.evict_before(4)
    assert sorted(w.iters) == [4, 5, 6, 7]


def test_three_file_format_round_trip(tmp_path):
    w = SampleWindow(device="cpu")
    o, p, z = _iters([50])[1]
    w.add_iteration(9, o, p, z)
    base = w.save_iteration(9, str(tmp_path), "run")
    assert base.endswith("iteration-0009")
    assert torch.equal(torch.load(base + "-data.pkl", weights_only=False), o)        # what Coach.train's add_tensor_dataset loads
    w2 = SampleWindow(device="cpu")
    assert w2.load_iteration(9, str(tmp_path), "run") and not w2.load_iteration(10, str(tmp_path), "run")
    for a, b in zip(w2.iters[9], (o, p, z)):
        assert torch.equal(a, b)


def test_losses_and_training_consume_the_loader():
    """NNetWrapper.train's loop body (NNetWrapper.py:131-165) over a WindowLoader."""
    from azb200 import nnet as aznet
    torch.manual_seed(0)
    net = aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS)
    opt = torch.optim.SGD(net.parameters(), lr=0.01)
    its = _iters([96, 64])
    for k in its:                                                # proper distributions as targets
        o, p, z = its[k]
        its[k] = (o, p / p.sum(1, keepdim=True), z / z.sum(1, keepdim=True))
    w = SampleWindow(device="cpu")
    for i, t in its.items():
        w.add_iteration(i, *t)
    args = dict(train_batch_size=32, minTrainHistoryWindow=4, maxTrainHistoryWindow=20, trainHistoryIncrementIters=2)
    loader, used = w.loader(2, args)
    assert used == [1, 2] and len(loader) == 5
    net.train()
    losses = []
    for boards, tp, tv in loader:
        out_pi, out_v = net(boards)
        l = loss_pi(tp, out_pi) + loss_v(tv, out_v, 1.0)
        want = -(tp * out_pi).sum() / len(boards) - (tv * out_v).sum() / len(boards)
        assert torch.allclose(l, want)
        opt.zero_grad(); l.backward(); opt.step()
        losses.append(float(l))
    assert np.isfinite(losses).all()


@pytest.mark.gpu
def test_window_fills_from_the_engine_on_the_device():
    """add_from_engine drains the engine's sample ring device-to-device; same samples as the host drain."""
    from azb200 import SelfPlayEngine, default_temp_scaling, temp_table
    kw = dict(game="connect4", num_games=64, rng="philox", seed=3, temps=temp_table(default_temp_scaling, 1, 42),
              max_sims_per_move=8)
    a, b = SelfPlayEngine(**kw), SelfPlayEngine(**kw)
    w = SampleWindow(device="cuda")
    host = [np.zeros((0, 4, 6, 7), np.float32), np.zeros((0, 7), np.float32), np.zeros((0, 3), np.float32)]
    for _ in range(30):
        for e in (a, b):
            e.warmup_sims(8); e.play_moves(False)
        w.add_from_engine(1, a)
        o, p, z, _ = b.drain_samples()
        host = [np.concatenate([x, y]) for x, y in zip(host, (o, p, z))]
    assert len(host[0]) > 100
    for x, y in zip(w.iters[1], host):
        assert np.array_equal(x.cpu().numpy(), y)
    loader, used = w.loader(1, dict(train_batch_size=128))
    boards, pis, vs = next(iter(loader))
    assert boards.is_cuda and boards.shape == (128, 4, 6, 7) and pis.shape == (128, 7) and vs.shape == (128, 3)
