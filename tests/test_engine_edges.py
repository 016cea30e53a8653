"""Edge cases of the engine through the C ABI: ragged batch sizes (fewer games than one warp holds, sizes that leave
a partially filled last warp / CTA, slot sub-ranges), and the error behaviour the reference has at the same places --
ValueError for a move that is no child (MCTS.pyx:195), FloatingPointError for a zero / NaN prior sum
(np.seterr(all='raise'), MCTS.pyx:23) -- plus the engine's own limits (node pool, sample ring, fed noise table)."""
import numpy as np
import pytest
import torch

import _orc
from _fakenn import FakeNN
from _lockstep import assert_queues_equal, assert_traces_equal, run_trace

pytestmark = pytest.mark.gpu

C4_TEMPS = _orc.temp_table(_orc.default_temp_scaling, 1, 42)


@pytest.mark.parametrize("game,B", [("connect4", 1), ("connect4", 3), ("connect4", 37), ("brandubh", 1), ("brandubh", 5)])
def test_ragged_batch_sizes_bit_exact(game, B):
    """B = 1, B smaller than the games one warp holds, B that leaves the last warp / CTA partially filled."""
    from _engine_agent import EngineAgent
    tafl = game == "brandubh"
    gid = _orc.GAME_BRANDUBH if tafl else _orc.GAME_CONNECT4
    sims, rounds = (8, 30) if tafl else (20, 50)
    kw = dict(temps=(_orc.temp_table(_orc.default_temp_scaling, 1, None) if tafl else C4_TEMPS), add_root_temp=True)
    nn = FakeNN(int(np.prod((5, 7, 7) if tafl else (4, 6, 7))), 588 if tafl else 7, seed=B)
    orc = _orc.OracleAgent(gid, B, rng_mode=_orc.RNG_PHILOX, seed=3, game_id_base=11, **kw)
    eng = EngineAgent(game, B, rng="philox", seed=3, game_id_base=11, max_sims_per_move=sims, **kw)
    to, te = run_trace(orc, nn, rounds, sims, keep_obs=True), run_trace(eng, nn, rounds, sims, keep_obs=True)
    assert_traces_equal(to, te, f"{game} B={B}")
    assert_queues_equal(orc, eng, f"{game} B={B}")


def test_slot_subranges_cover_the_batch_like_one_launch():
    """select / expand_backup on [first, first+count) pieces (how cohorts split a batch) = one launch over all slots."""
    from azb200 import SelfPlayEngine
    B, sims = 50, 16
    out = []
    for pieces in ([(0, 50)], [(0, 7), (7, 1), (8, 29), (37, 13)]):
        eng = SelfPlayEngine(game="connect4", num_games=B, rng="philox", seed=4, temps=C4_TEMPS, add_root_temp=True,
                             max_sims_per_move=sims)
        nn = FakeNN(4 * 6 * 7, 7, seed=1)
        counts = []
        for _ in range(12):
            for _ in range(sims):
                for f, c in pieces:
                    eng.select(f, c)
                p, v = nn(eng.obs.cpu().numpy())
                eng.policy.copy_(torch.from_numpy(p)); eng.value.copy_(torch.from_numpy(v))
                for f, c in pieces:
                    eng.expand_backup(f, c)
            counts.append(eng.root_counts().copy())
            eng.play_moves(False)
        eng.check_errors()
        out.append((np.stack(counts), eng.turns().copy(), eng.stats()["sum_children"]))
        eng.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


def test_bad_configurations_are_refused():
    from azb200 import SelfPlayEngine
    from azb200._capi import AzbError
    for kw in (dict(num_games=0), dict(num_games=-4), dict(num_games=3, arena=True),
               dict(num_games=4, arena=True, add_root_noise=True), dict(num_games=4, game=99),
               dict(num_games=4, max_nodes_per_game=3), dict(num_games=4, lanes_per_game=5),
               dict(num_games=4, root_policy_temp=0.0), dict(num_games=4, device=99)):
        with pytest.raises(AzbError) as ei:
            SelfPlayEngine(**kw)
        assert ei.value.status == -1, kw                        # AZB_ERR_BAD_CONFIG
    eng = SelfPlayEngine(num_games=4)
    for f, c in ((-1, 2), (0, 5), (4, 1), (3, 2)):
        with pytest.raises(AzbError) as ei:
            eng.select(f, c)
        assert ei.value.status == -7                            # AZB_ERR_BAD_ARGUMENT
    eng.close()


def _round(eng, sims, policy=None, value=None):
    for _ in range(sims):
        eng.select()
        eng.policy.copy_(policy if policy is not None else torch.full_like(eng.policy, 1.0 / eng.A))
        eng.value.copy_(value if value is not None else torch.full_like(eng.value, 1.0 / 3))
        eng.expand_backup()


@pytest.mark.parametrize("bad", ["nan", "zero"])
def test_zero_or_nan_prior_sum_is_a_floating_point_error(bad):
    """process_results divides the masked policy by its sum under np.seterr(all='raise') (MCTS.pyx:23,247-249)."""
    from azb200 import SelfPlayEngine
    from azb200._capi import AzbError
    eng = SelfPlayEngine(num_games=8, max_sims_per_move=4)
    pol = torch.full_like(eng.policy, float("nan") if bad == "nan" else 0.0)
    _round(eng, 2, policy=pol)
    with pytest.raises(AzbError) as ei:
        eng.check_errors()
    assert ei.value.status == -5                                # AZB_ERR_FLOATING_POINT
    eng.close()


def test_node_pool_exhaustion_is_reported_not_overrun():
    from azb200 import SelfPlayEngine
    from azb200._capi import AzbError
    eng = SelfPlayEngine(num_games=8, max_sims_per_move=50, max_nodes_per_game=64)
    _round(eng, 50)
    with pytest.raises(AzbError) as ei:
        eng.check_errors()
    assert ei.value.status == -3                                # AZB_ERR_POOL_EXHAUSTED
    eng.close()


def test_sample_ring_overflow_is_reported():
    from azb200 import SelfPlayEngine
    from azb200._capi import AzbError
    eng = SelfPlayEngine(num_games=16, max_sims_per_move=2, sample_capacity=8, temps=C4_TEMPS)
    with pytest.raises(AzbError) as ei:
        for _ in range(45):                                     # every game ends within 42 plies
            _round(eng, 2)
            eng.play_moves(False)
        eng.check_errors()
    assert ei.value.status == -6                                # AZB_ERR_SAMPLE_OVERFLOW
    eng.close()


def test_fed_root_noise_underrun_is_reported():
    from azb200 import SelfPlayEngine
    from azb200._capi import AzbError
    # the tree is rebuilt after every move (mctsResetThreshold = 1), so every move expands a root and takes a noise row
    eng = SelfPlayEngine(num_games=4, max_sims_per_move=3, add_root_noise=True, temps=C4_TEMPS, mcts_reset_threshold=1)
    eng.set_root_noise(np.full((4, 2, 7), 1.0 / 7, dtype=np.float32))       # two root expansions per game, then dry
    with pytest.raises(AzbError) as ei:
        for _ in range(6):
            _round(eng, 3)
            eng.play_moves(False)
        eng.check_errors()
    assert ei.value.status == -8                                # AZB_ERR_NOISE_UNDERRUN
    eng.close()


def test_draining_empty_queues():
    from azb200 import SelfPlayEngine
    eng = SelfPlayEngine(num_games=4, max_sims_per_move=2)
    obs, pi, z, slot = eng.drain_samples()
    assert obs.shape == (0, 4, 6, 7) and pi.shape == (0, 7) and z.shape == (0, 3) and slot.shape == (0,)
    s, t, w = eng.drain_results()
    assert len(s) == len(t) == len(w) == 0 and eng.games_played() == 0 and eng.sample_count() == 0
    eng.close()
