"""Config 5 in miniature: GpuCoach.learn -- warmup self-play, training on the device window, arena gating, a second
iteration with the network in the loop -- runs end to end on one GPU and keeps the reference's bookkeeping."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _C4:
    __module__ = "alphazero.envs.connect4.connect4"

    @staticmethod
    def observation_size():
        return (4, 6, 7)

    @staticmethod
    def action_size():
        return 7

    @staticmethod
    def num_players():
        return 2

    @staticmethod
    def max_turns():
        return 42


def test_two_iterations_of_the_full_loop():
    from azb200.loop import GpuCoach, winrate_of_first
    assert winrate_of_first([6, 2], 2, True) == pytest.approx((6 + 1) / 10)          # draws count half
    assert winrate_of_first([6, 2], 2, False) == pytest.approx(6 / 8)
    coach = GpuCoach(_C4, dict(numIters=2, numWarmupIters=1, gamesPerIteration=96, process_batch_size=64, numMCTSSims=12,
                               numFastSims=6, numWarmupSims=5, probFastSim=0.5, train_batch_size=64, arenaCompare=24,
                               min_next_model_winrate=0.0), seed=1)
    hist = coach.learn()
    assert [h["iteration"] for h in hist] == [1, 2]
    assert hist[0]["warmup"] and hist[0]["samples"] > 96 * 7 * 2 * 0.5 and hist[0]["train_steps"] == hist[0]["samples"] // 64
    assert all(np.isfinite([h["loss_pi"], h["loss_v"]]).all() for h in hist)
    for h in hist:
        assert sum(h["arena_wins"]) + h["arena_draws"] == 24 and h["accepted"]       # threshold 0: always accepted
        assert sum(h["game_results"][0]) + h["game_results"][1] >= 96
    assert not hist[1]["warmup"] and hist[1]["window"] == [1, 2] and coach.self_play_iter == 2
    a = {k: v.clone() for k, v in coach.self_play_net.nnet.state_dict().items()}
    b = coach.train_net.nnet.state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_graphed_training_step_equals_the_eager_loop():
    """train_steps replays the captured step after three eager ones; ragged last batches of an epoch stay eager.
    Same batches, same order -> the same parameters and losses as the all-eager loop (up to cuDNN's non-deterministic
    summation order in the weight gradients)."""
    from azb200 import nnet as aznet
    from azb200.loop import train_steps
    from azb200.samples import WindowLoader
    g = torch.Generator().manual_seed(2)
    n = 5 * 64 + 20
    obs = torch.rand(n, 4, 6, 7, generator=g).cuda()
    pi = torch.softmax(torch.randn(n, 7, generator=g), 1).cuda()
    z = torch.softmax(torch.randn(n, 3, generator=g), 1).cuda()
    out = []
    for graph_after in (3, None):
        torch.manual_seed(0)
        w = aznet.NNetWrapper(nnet=aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS).cuda(), cuda=True)
        opt = torch.optim.SGD(w.nnet.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
        torch.manual_seed(5)
        losses = train_steps(w, opt, WindowLoader((obs, pi, z), 64), 15, 1.5, graph_after=graph_after)   # 2.5 epochs
        assert (len(w.__dict__.get("_graphed_steps", {})) == 1) == (graph_after is not None)
        out.append((losses, torch.cat([p.detach().reshape(-1) for p in w.nnet.parameters()]),
                    torch.cat([b.detach().reshape(-1).float() for b in w.nnet.buffers()])))
    # two all-eager runs differ by ~2e-4 in the parameters after 15 steps (cuDNN's weight gradients are not
    # deterministic); the graphed run sits in the same band
    assert out[0][0] == pytest.approx(out[1][0], rel=1e-4)
    assert torch.allclose(out[0][1], out[1][1], atol=2e-3) and torch.allclose(out[0][2], out[1][2], atol=2e-3)


def test_folding_a_network_leaves_it_untouched_so_the_captured_step_stays_valid():
    """The fused evaluator folds a private copy of the model: storage, device, dtype and mode of the caller's network
    are unchanged (moving it through .cpu() / .double() would re-allocate every parameter under a captured training
    step -- the arena evaluates train_net between two training phases)."""
    from azb200 import nnet as aznet
    from azb200.fused_nn import FusedResNetEvaluator
    from azb200.loop import train_steps
    from azb200.samples import WindowLoader
    g = torch.Generator().manual_seed(3)
    n = 8 * 64
    obs = (torch.rand(n, 4, 6, 7, generator=g) > 0.5).float().cuda()
    pi = torch.softmax(torch.randn(n, 7, generator=g), 1).cuda()
    z = torch.softmax(torch.randn(n, 3, generator=g), 1).cuda()
    torch.manual_seed(0)
    w = aznet.NNetWrapper(nnet=aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS).cuda(), cuda=True)
    opt = torch.optim.SGD(w.nnet.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)
    train_steps(w, opt, WindowLoader((obs, pi, z), 64), 8, 1.5)
    assert len(w._graphed_steps) == 1
    w.nnet.train()
    before = [(p.data_ptr(), p.dtype, p.device) for p in list(w.nnet.parameters()) + list(w.nnet.buffers())]
    pol, val = torch.zeros(64, 7, device="cuda"), torch.zeros(64, 3, device="cuda")
    ev = FusedResNetEvaluator(w.nnet, obs[:64].contiguous(), pol, val)
    ev()
    torch.cuda.synchronize()
    assert before == [(p.data_ptr(), p.dtype, p.device) for p in list(w.nnet.parameters()) + list(w.nnet.buffers())]
    assert w.nnet.training and torch.isfinite(pol).all()
    lp, lv = train_steps(w, opt, WindowLoader((obs, pi, z), 64), 8, 1.5)          # replays the graph captured above
    assert np.isfinite([lp, lv]).all() and all(torch.isfinite(p).all() for p in w.nnet.parameters())
