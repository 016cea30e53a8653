"""azb200.mcts.MCTS (single-tree API on the engine) against the compiled reference's MCTS class (oracle/_ref):
search / raw_search / counts / probs / value / best_action / update_root over a played game, same RNG stream."""
import numpy as np
import pytest

import _refdriver
from _fakenn import FakeNN

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _refdriver.available(), reason="oracle/_ref not built")]


def _args():
    _, _, dotdict, _ = _refdriver._import_ref()
    return dotdict(dict(cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1, min_discount=1,
                        _num_players=2, numMCTSSims=40))


def _single(net):
    def nn(obs):
        p, v = net(np.asarray(obs, dtype=np.float32)[None])
        return p[0], v[0]
    return nn


@pytest.mark.parametrize("game,raw", [("connect4", False), ("connect4", True), ("brandubh", False)])
def test_search_api_matches_reference(game, raw):
    from azb200.mcts import MCTS
    ref_mcts, _, _, _ = _refdriver._import_ref()
    Game = _refdriver.game_class(game)
    args = _args()
    c4 = game == "connect4"
    net = FakeNN(4 * 6 * 7 if c4 else 5 * 7 * 7, 7 if c4 else 588, seed=8, sharp=3.0 if c4 else 1.0)
    nn = _single(net)
    sims, moves, seed = (25, 14, 7) if c4 else (10, 8, 9)
    np.random.seed(seed)
    ref = ref_mcts.MCTS(args)
    mine = MCTS(args, rng="mt19937", seed=seed)
    gs = Game()
    rs = np.random.RandomState(1)
    for mv in range(moves):
        state = np.random.get_state()                       # the reference draws its shuffles from the global stream
        if raw:
            ref.raw_search(gs, sims, False, False)
        else:
            ref.search(gs, nn, sims, False, False)
        after = np.random.get_state()
        if raw:
            mine.raw_search(gs, sims, False, False)
        else:
            mine.search(gs, nn, sims, False, False)
        np.random.set_state(after)
        want = np.asarray(ref.counts(gs))
        got = mine.counts(gs)
        assert np.array_equal(want, got), (mv, want, got)
        assert mine.best_action(gs) == ref.best_action(gs)
        assert np.array_equal(mine.probs(gs, 1.0), np.asarray(ref.probs(gs, 1.0)))
        assert np.array_equal(mine.probs(gs, 0), np.asarray(ref.probs(gs, 0)))
        assert mine.value() == pytest.approx(float(ref.value()), abs=0) and mine.value(True) == pytest.approx(float(ref.value(True)), rel=1e-6)
        assert mine.max_depth == ref.max_depth
        legal = np.flatnonzero(np.asarray(gs.valid_moves()))
        a = int(rs.choice(legal)) if mv % 3 == 2 else int(np.argmax(want))     # sometimes a rarely visited move
        ref.update_root(gs, a)
        mine.update_root(gs, a)
        gs.play_action(a)
        if np.asarray(gs.win_state()).any():
            break
    with pytest.raises(ValueError):
        full = Game()
        mine.reset()
        mine.search(full, nn, 2, False, False)
        mine.update_root(full, 10 ** 6)
