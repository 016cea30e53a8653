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
