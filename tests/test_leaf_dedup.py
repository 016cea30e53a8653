"""Leaf de-duplication (azb_set_leaf_dedup, include/azb200.h): games whose leaves have the same observation share one
network evaluation.  The engine must stay bit-equal to the oracle -- which, like the reference's NN server
(Coach.py:337-342), evaluates every row -- when ONLY the rows the engine lists are answered and every other row of the
policy / value tensors is poisoned."""
import numpy as np
import pytest
import torch

import _orc
from _fakenn import FakeNN
from _lockstep import assert_queues_equal, assert_traces_equal, run_trace

pytestmark = pytest.mark.gpu

C4_TEMPS = _orc.temp_table(_orc.default_temp_scaling, 1, 42)


def _dedup_agent(*args, **kw):
    from _engine_agent import EngineAgent

    class DedupEngineAgent(EngineAgent):
        """answers only the rows azb_nn_rows_ptr lists after the last select; NaN everywhere else"""

        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.eng.set_leaf_dedup(True)
            self.listed = 0
            self.rows_seen = []

        def processBatch(self, policy, value):
            e = self.eng
            n = int(e._wrap(e.nn_count_ptr(), (1,), "<i4").item())
            rows = e.nn_rows[:n].cpu().numpy()
            assert len(set(rows.tolist())) == n and (n == 0 or (rows.min() >= 0 and rows.max() < self.B))
            obs = e.obs.cpu().numpy().reshape(self.B, -1)
            # the listed rows have pairwise different observations
            assert len({obs[r].tobytes() for r in rows}) == n
            self.listed += n
            self.rows_seen.append(n)
            p = np.full(policy.shape, np.nan, dtype=np.float32)
            v = np.full(value.shape, np.nan, dtype=np.float32)
            p[rows] = policy[rows]; v[rows] = value[rows]
            super().processBatch(p, v)

    return DedupEngineAgent(*args, **kw)


@pytest.mark.parametrize("game", ["connect4", "brandubh"])
@pytest.mark.parametrize("fused", [False, True])
def test_engine_with_dedup_equals_oracle(game, fused):
    """fresh games (the first move-rounds are almost all duplicates) through to finished and restarted games; separate
    launches and the fused expand/backup + select launch"""
    c4 = game == "connect4"
    B, sims, rounds = (96, 11, 50) if c4 else (6, 7, 25)
    obs_n, A = (4 * 6 * 7, 7) if c4 else (5 * 7 * 7, 588)
    temps = C4_TEMPS if c4 else _orc.temp_table(_orc.default_temp_scaling, 1, None)
    nn = FakeNN(obs_n, A, seed=21, sharp=3.0 if c4 else 1.0)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4 if c4 else _orc.GAME_BRANDUBH, B, rng_mode=_orc.RNG_PHILOX, seed=4,
                           add_root_temp=True, temps=temps)
    eng = _dedup_agent(game, B, rng="philox", seed=4, add_root_temp=True, temps=temps, max_sims_per_move=sims,
                       fused_step_sims=sims if fused else 0)
    assert_traces_equal(run_trace(orc, nn, rounds, sims, keep_obs=True), run_trace(eng, nn, rounds, sims, keep_obs=True), game)
    assert_queues_equal(orc, eng, game)
    so, se = orc.stats(), eng.stats()
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "moves"):
        assert so[k] == se[k], (k, so[k], se[k])
    dups = eng.eng.duplicate_leaves()
    assert eng.listed + dups == se["sims"] - se["terminal_leaves"]
    assert dups > 0 and eng.rows_seen[0] == 1          # every game's first leaf is the empty board's root
    # switching it off again: every non-terminal leaf is listed
    eng.eng.set_leaf_dedup(False)
    t0 = eng.eng.stats()["terminal_leaves"]
    eng.eng.select()
    n = int(eng.eng._wrap(eng.eng.nn_count_ptr(), (1,), "<i4").item())
    assert n == B - (eng.eng.stats()["terminal_leaves"] - t0)


def test_mt19937_parity_mode_with_dedup():
    """NumPy-stream parity mode (per-slot MT19937), root noise fed from the host"""
    B, sims, rounds = 48, 9, 30
    seeds = list(range(500, 500 + B))
    rs = np.random.RandomState(3)
    noise = rs.dirichlet([1.0] * 7, size=(B, 64)).astype(np.float32)
    nn = FakeNN(4 * 6 * 7, 7, seed=5)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, rng_mode=_orc.RNG_MT19937, mt_seeds=seeds, add_root_noise=True,
                           add_root_temp=True, temps=C4_TEMPS)
    eng = _dedup_agent("connect4", B, rng="mt19937", mt_seeds=seeds, add_root_noise=True, add_root_temp=True, temps=C4_TEMPS,
                       max_sims_per_move=sims)
    orc.set_root_noise(noise); eng.set_root_noise(noise)
    assert_traces_equal(run_trace(orc, nn, rounds, sims), run_trace(eng, nn, rounds, sims), "connect4 mt19937")
    assert_queues_equal(orc, eng, "connect4 mt19937")
    assert eng.eng.duplicate_leaves() > 0


def test_dedup_refused_in_arena_mode():
    from azb200 import SelfPlayEngine
    eng = SelfPlayEngine(game="connect4", num_games=8, rng="philox", seed=1, arena=True, temps=np.array([0.25]),
                         add_root_noise=False, add_root_temp=False)
    with pytest.raises(Exception):
        eng.set_leaf_dedup(True)
    eng.close()


@pytest.mark.parametrize("game,B,sims", [("connect4", 2048, 30), ("brandubh", 256, 20)])
def test_device_selfplay_with_and_without_dedup_is_identical(game, B, sims):
    """DeviceSelfPlay with the tcgen05 evaluator, the round replayed as a CUDA graph (the graph repeats its epochs, the
    table is cleared by the memset node of play_moves): examples, results, root counts and statistics are bit-identical
    with and without leaf de-duplication, from fresh games on."""
    from azb200 import SelfPlayEngine, default_temp_scaling, temp_table
    from azb200 import nnet as aznet
    from azb200.selfplay import DeviceSelfPlay
    tafl = game == "brandubh"
    OBS, A = ((5, 7, 7), 588) if tafl else ((4, 6, 7), 7)
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = aznet.ResNet(OBS, A, 3, **(aznet.BRANDUBH_TRAIN_NET_ARGS if tafl else aznet.DEFAULT_NET_ARGS)).to(dev).eval()
    outs = []
    for dedup in (False, True):
        eng = SelfPlayEngine(game=game, num_games=B, rng="philox", seed=9, add_root_noise=True, add_root_temp=True,
                             symmetric_samples=True, max_sims_per_move=sims, sample_capacity=400000,
                             temps=temp_table(default_temp_scaling, 1, None if tafl else 42))
        drv = DeviceSelfPlay(eng, model, dedup=dedup)
        assert drv.dedup == dedup and drv.round_graph
        counts, samples, results = [], [], []
        for r in range(7):
            drv.run_round(sims)
            torch.cuda.synchronize()
            eng.check_errors()
            counts.append(eng.root_counts().copy())
            samples.append(eng.drain_samples())
            results.append(eng.drain_results())
        outs.append((counts, samples, results, eng.stats(), eng.duplicate_leaves()))
        eng.close()
    (c0, s0, r0, st0, d0), (c1, s1, r1, st1, d1) = outs
    for a, b in zip(c0, c1):
        assert np.array_equal(a, b)
    for a, b in zip(s0, s1):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for a, b in zip(r0, r1):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "moves", "samples", "results"):
        assert st0[k] == st1[k], k
    # fresh games: Connect4 is mostly duplicates for the first moves; brandubh (40 first moves) measured 27 %
    assert d0 == 0 and d1 > (0.3 if game == "connect4" else 0.1) * (st1["sims"] - st1["terminal_leaves"])
