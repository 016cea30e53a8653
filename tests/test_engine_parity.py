"""GPU parity tests proper: the CUDA engine, called through the C ABI, against
the CPU oracle (oracle/liborc.so) on the same seeded inputs.  Bit-exact: visit
counts, sampled actions, turns, leaf observations, emitted samples (obs, pi, z,
order) and game results."""
import numpy as np
import pytest

import _orc
from _fakenn import FakeNN
from _lockstep import assert_queues_equal, assert_traces_equal, run_trace

pytestmark = pytest.mark.gpu

C4_TEMPS = _orc.temp_table(_orc.default_temp_scaling, 1, 42)


def _pair(B, rng, seeds=None, noise=None, **kw):
    from _engine_agent import EngineAgent
    okw = {k: v for k, v in kw.items() if k not in ("max_sims_per_move", "max_nodes_per_game")}
    ekw = dict(kw)
    if rng == "mt19937":
        seeds = list(range(100, 100 + B)) if seeds is None else seeds
        orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, rng_mode=_orc.RNG_MT19937, mt_seeds=seeds, **okw)
        eng = EngineAgent("connect4", B, rng="mt19937", mt_seeds=seeds, **ekw)
    else:
        orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, rng_mode=_orc.RNG_PHILOX, seed=77, game_id_base=5, **okw)
        eng = EngineAgent("connect4", B, rng="philox", seed=77, game_id_base=5, **ekw)
    if noise is not None:
        orc.set_root_noise(noise)
        eng.set_root_noise(noise)
    return orc, eng


@pytest.mark.parametrize("rng", ["mt19937", "philox"])
@pytest.mark.parametrize("mode", ["warmup", "nn"])
def test_connect4_selfplay_bit_exact(rng, mode):
    B, rounds, sims = 16, 70, 25
    nn = FakeNN(4 * 6 * 7, 7) if mode == "nn" else None
    orc, eng = _pair(B, rng, temps=C4_TEMPS, add_root_temp=True, max_sims_per_move=sims)
    to = run_trace(orc, nn, rounds, sims, keep_obs=True)
    te = run_trace(eng, nn, rounds, sims, keep_obs=True)
    assert_traces_equal(to, te, f"{rng}/{mode}")
    assert_queues_equal(orc, eng, f"{rng}/{mode}")
    so, se = orc.stats(), eng.stats()
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "games_played",
              "results", "samples", "moves"):
        assert so[k] == se[k], (k, so[k], se[k])
    assert so["games_played"] > 0 and so["samples"] > 0


def test_connect4_root_noise_and_tree_values():
    B, rounds, sims = 8, 30, 40
    rs = np.random.RandomState(5)
    noise = rs.dirichlet([10.83 / 7] * 7, size=(B, 16)).astype(np.float32)
    nn = FakeNN(4 * 6 * 7, 7, seed=9)
    orc, eng = _pair(B, "mt19937", noise=noise, temps=C4_TEMPS, add_root_temp=True, add_root_noise=True,
                     max_sims_per_move=sims)
    to = run_trace(orc, nn, rounds, sims)
    te = run_trace(eng, nn, rounds, sims)
    assert_traces_equal(to, te, "noise")
    assert_queues_equal(orc, eng, "noise")


def test_connect4_fast_moves_quota_and_reset_threshold():
    B, sims, quota = 12, 12, 20
    nn = FakeNN(4 * 6 * 7, 7, seed=3)
    orc, eng = _pair(B, "philox", temps=C4_TEMPS, games_per_iteration=quota, mcts_reset_threshold=5,
                     symmetric_samples=False, max_sims_per_move=sims)
    pat = [0, 1, 1, 0, 1]
    # SelfPlayAgent.run stops once games_played reaches gamesPerIteration (SelfPlayAgent.pyx:82)
    to = run_trace(orc, nn, 400, sims, fast_pattern=pat, until_games=quota)
    te = run_trace(eng, nn, 400, sims, fast_pattern=pat, until_games=quota)
    assert len(to) == len(te) < 400
    assert_traces_equal(to, te, "quota")
    assert_queues_equal(orc, eng, "quota")
    assert orc.stats()["games_played"] == eng.stats()["games_played"] == quota
    assert eng.stats()["results"] >= quota


def test_tree_dump_matches_oracle_reference_values():
    """Node fields (n, q, v, p, player, e) of a whole tree against the compiled
    reference when it is available, else structure sanity only."""
    import _refdriver
    if not _refdriver.available():
        pytest.skip("oracle/_ref not built")
    from _engine_agent import EngineAgent
    B, sims = 3, 60
    seeds = [7, 8, 9]
    nn = FakeNN(4 * 6 * 7, 7, seed=11)
    ref = _refdriver.RefAgent("connect4", B, mt_seeds=seeds, add_root_temp=True, det_pow=True)
    eng = EngineAgent("connect4", B, rng="mt19937", mt_seeds=seeds, temps=C4_TEMPS, add_root_temp=True,
                      max_sims_per_move=sims)
    for agent in (ref, eng):
        run_trace(agent, nn, 3, sims)
        for _ in range(sims):
            obs = agent.generateBatch()
            agent.processBatch(*nn(obs))
    for s in range(B):
        a, b = ref.tree_dump(s), eng.tree_dump(s)
        # the reference materialises (never used) children of terminal leaves; drop them
        keep, skip_depth = [], None
        for row in a:
            if skip_depth is not None and row[0] > skip_depth:
                continue
            skip_depth = None
            keep.append(row)
            if row[7] or row[8] or row[9]:
                skip_depth = row[0]
        a = np.asarray(keep)
        assert a.shape == b.shape, (a.shape, b.shape)
        assert np.array_equal(a.astype(np.float32), b.astype(np.float32))


TAFL_TEMPS = _orc.temp_table(_orc.default_temp_scaling, 1, None)


def _tafl_pair(B, rng, noise=None, **kw):
    from _engine_agent import EngineAgent
    okw = {k: v for k, v in kw.items() if k not in ("max_sims_per_move", "max_nodes_per_game")}
    if rng == "mt19937":
        seeds = list(range(200, 200 + B))
        orc = _orc.OracleAgent(_orc.GAME_BRANDUBH, B, rng_mode=_orc.RNG_MT19937, mt_seeds=seeds, **okw)
        eng = EngineAgent("brandubh", B, rng="mt19937", mt_seeds=seeds, **kw)
    else:
        orc = _orc.OracleAgent(_orc.GAME_BRANDUBH, B, rng_mode=_orc.RNG_PHILOX, seed=9, game_id_base=3, **okw)
        eng = EngineAgent("brandubh", B, rng="philox", seed=9, game_id_base=3, **kw)
    if noise is not None:
        orc.set_root_noise(noise)
        eng.set_root_noise(noise)
    return orc, eng


@pytest.mark.parametrize("rng", ["mt19937", "philox"])
@pytest.mark.parametrize("mode", ["warmup", "nn"])
def test_brandubh_selfplay_bit_exact(rng, mode):
    B, rounds, sims = 6, 120, 16
    nn = FakeNN(5 * 7 * 7, 588, seed=21, sharp=1.0) if mode == "nn" else None
    orc, eng = _tafl_pair(B, rng, temps=TAFL_TEMPS, add_root_temp=True, max_sims_per_move=sims)
    to = run_trace(orc, nn, rounds, sims, keep_obs=True)
    te = run_trace(eng, nn, rounds, sims, keep_obs=True)
    assert_traces_equal(to, te, f"tafl {rng}/{mode}")
    assert_queues_equal(orc, eng, f"tafl {rng}/{mode}")
    so, se = orc.stats(), eng.stats()
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "games_played",
              "results", "samples", "moves"):
        assert so[k] == se[k], (k, so[k], se[k])
    assert so["games_played"] > 0 and so["samples"] > 0
    assert np.array_equal(orc.boards(), eng.boards())


def test_brandubh_root_noise_deep_search():
    B, rounds, sims = 3, 12, 150
    noise = np.random.RandomState(8).dirichlet([10.83 / 40] * 96, size=(B, 4)).astype(np.float32)
    nn = FakeNN(5 * 7 * 7, 588, seed=22, sharp=2.0)
    orc, eng = _tafl_pair(B, "mt19937", noise=noise, temps=TAFL_TEMPS, add_root_temp=True, add_root_noise=True,
                          max_sims_per_move=sims)
    assert_traces_equal(run_trace(orc, nn, rounds, sims), run_trace(eng, nn, rounds, sims), "tafl noise")
    assert np.array_equal(orc.boards(), eng.boards())


@pytest.mark.parametrize("game,rng,p2i,reset", [("connect4", "mt19937", [0, 1], None), ("connect4", "philox", [1, 0], 5),
                                                ("brandubh", "philox", [0, 1], None), ("hnefatafl", "mt19937", [1, 0], None)])
def test_arena_mode_equals_oracle(game, rng, p2i, reset):
    """SelfPlayAgent(_is_arena=True) on the engine (azb_config.arena): visit counts of the searching tree, sampled
    actions, leaf observations, results and the quota, bit-equal to the oracle's arena mode (itself pinned against the
    compiled reference in test_oracle_vs_ref.py)."""
    from _engine_agent import ArenaEngineAgent
    from _fakenn import ArenaNN
    c4 = game == "connect4"
    B, sims, quota = (6, 9, 14) if c4 else (3, 6, 1 << 40)
    obs_n, A = (4 * 6 * 7, 7) if c4 else (5 * 11 * 11, 2420) if game == "hnefatafl" else (5 * 7 * 7, 588)
    gid = {"connect4": _orc.GAME_CONNECT4, "brandubh": _orc.GAME_BRANDUBH, "hnefatafl": _orc.GAME_HNEFATAFL}[game]
    seeds = list(range(31, 31 + B))
    nets = [FakeNN(obs_n, A, seed=70, sharp=3.0 if c4 else 1.0), FakeNN(obs_n, A, seed=71, sharp=3.0 if c4 else 1.0)]
    kw = dict(mt_seeds=seeds) if rng == "mt19937" else dict(seed=9)
    orc = _orc.OracleAgent(gid, B, arena=True, arena_temp=0.25,
                           rng_mode=_orc.RNG_MT19937 if rng == "mt19937" else _orc.RNG_PHILOX, player_to_index=p2i,
                           mcts_reset_threshold=reset or 0, games_per_iteration=quota, **kw)
    eng = ArenaEngineAgent(game, B, rng=rng, arena_temp=0.25, player_to_index=p2i, mcts_reset_threshold=reset,
                           games_per_iteration=quota, max_sims_per_move=sims, **kw)
    rounds = 200 if c4 else 30
    until = quota if c4 else None
    # stop at the first finished-beyond-quota game: the worker loop's exit condition (SelfPlayAgent.pyx:82)
    ta = run_trace(orc, ArenaNN(orc, nets), rounds, sims, keep_obs=True, until_games=until)
    tb = run_trace(eng, ArenaNN(eng, nets), rounds, sims, keep_obs=True, until_games=until)
    assert_traces_equal(ta, tb, f"arena {game} {rng}")
    for name, x, y in zip(("slot", "turns", "win"), orc.results(), eng.results()):
        assert np.array_equal(x, y), f"arena results.{name} differ"
    so, se = orc.stats(), eng.stats()
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "games_played", "results", "moves"):
        assert so[k] == se[k], (k, so[k], se[k])
    assert se["samples"] == 0


@pytest.mark.parametrize("game", ["connect4", "brandubh"])
def test_fused_expand_select_launch_is_identical(game):
    """azb_expand_backup_select (processBatch of simulation k + generateBatch of k+1 in one launch) against the
    oracle: traces, samples, results and statistics as with the two separate calls."""
    from _engine_agent import EngineAgent
    c4 = game == "connect4"
    B, sims, rounds = (12, 11, 60) if c4 else (3, 7, 25)
    obs_n, A = (4 * 6 * 7, 7) if c4 else (5 * 7 * 7, 588)
    temps = C4_TEMPS if c4 else _orc.temp_table(_orc.default_temp_scaling, 1, None)
    nn = FakeNN(obs_n, A, seed=21, sharp=3.0 if c4 else 1.0)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4 if c4 else _orc.GAME_BRANDUBH, B, rng_mode=_orc.RNG_PHILOX, seed=4,
                           add_root_temp=True, temps=temps)
    eng = EngineAgent(game, B, rng="philox", seed=4, add_root_temp=True, temps=temps, max_sims_per_move=sims,
                      fused_step_sims=sims)
    assert_traces_equal(run_trace(orc, nn, rounds, sims, keep_obs=True), run_trace(eng, nn, rounds, sims, keep_obs=True), game)
    assert_queues_equal(orc, eng, game)
    so, se = orc.stats(), eng.stats()
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "moves"):
        assert so[k] == se[k], (k, so[k], se[k])


def test_select_lists_the_leaves_that_need_the_network():
    """azb_nn_rows_ptr / azb_nn_count_ptr after every select: exactly the live slots whose leaf is not terminal
    (the oracle's terminal_leaves statistic counts the complement), across move-rounds and counter parities."""
    from _engine_agent import EngineAgent
    B, sims, rounds = 64, 9, 30
    nn = FakeNN(4 * 6 * 7, 7, seed=33)
    eng = EngineAgent("connect4", B, rng="philox", seed=6, temps=C4_TEMPS, max_sims_per_move=sims)
    e = eng.eng
    total_listed, before = 0, e.stats()
    for r in range(rounds):
        for s in range(sims):
            t0 = e.stats()["terminal_leaves"]
            obs = eng.generateBatch()
            n = int(e._wrap(e.nn_count_ptr(), (1,), "<i4").item())
            rows = e.nn_rows[:n].cpu().numpy()
            term = e.stats()["terminal_leaves"] - t0
            assert n == B - term and len(set(rows.tolist())) == n and rows.min() >= 0 and rows.max() < B
            total_listed += n
            eng.processBatch(*nn(obs))
        eng.playMoves(False)
    st = e.stats()
    assert total_listed == (st["sims"] - before["sims"]) - (st["terminal_leaves"] - before["terminal_leaves"])
    assert st["terminal_leaves"] > 0


# ---- hnefatafl 11x11: 128-bit bitboards, 2420 actions ----------------------------------------------------
def _hnef_pair(B, rng, noise=None, arena=False, **kw):
    from _engine_agent import ArenaEngineAgent, EngineAgent
    okw = {k: v for k, v in kw.items() if k not in ("max_sims_per_move", "max_nodes_per_game")}
    mk = ArenaEngineAgent if arena else EngineAgent
    if rng == "mt19937":
        seeds = list(range(300, 300 + B))
        orc = _orc.OracleAgent(_orc.GAME_HNEFATAFL, B, rng_mode=_orc.RNG_MT19937, mt_seeds=seeds, arena=arena, **okw)
        eng = mk("hnefatafl", B, rng="mt19937", mt_seeds=seeds, **kw)
    else:
        orc = _orc.OracleAgent(_orc.GAME_HNEFATAFL, B, rng_mode=_orc.RNG_PHILOX, seed=11, game_id_base=5, arena=arena, **okw)
        eng = mk("hnefatafl", B, rng="philox", seed=11, game_id_base=5, **kw)
    if noise is not None:
        orc.set_root_noise(noise)
        eng.set_root_noise(noise)
    return orc, eng


@pytest.mark.gpu
@pytest.mark.parametrize("rng", ["mt19937", "philox"])
@pytest.mark.parametrize("mode", ["warmup", "nn"])
def test_hnefatafl_selfplay_bit_exact(rng, mode):
    """Engine == oracle (pinned to the compiled reference's hnefatafl env, tests/test_hnefatafl_rules.py): visit counts,
    actions, turns, leaf observations, the 8-fold symmetric samples with the 2420-entry permutation, results,
    statistics and boards -- games end (king's-side wins) inside the run."""
    B, rounds, sims = 4, 90, 12
    nn = FakeNN(5 * 11 * 11, 2420, seed=31, sharp=1.0) if mode == "nn" else None
    orc, eng = _hnef_pair(B, rng, temps=TAFL_TEMPS, add_root_temp=True, max_sims_per_move=sims)
    to = run_trace(orc, nn, rounds, sims, keep_obs=True)
    te = run_trace(eng, nn, rounds, sims, keep_obs=True)
    assert_traces_equal(to, te, f"hnefatafl {rng}/{mode}")
    assert_queues_equal(orc, eng, f"hnefatafl {rng}/{mode}")
    so, se = orc.stats(), eng.stats()
    for k in ("sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves", "games_played",
              "results", "samples", "moves"):
        assert so[k] == se[k], (k, so[k], se[k])
    assert np.array_equal(orc.boards(), eng.boards())


@pytest.mark.gpu
def test_hnefatafl_root_noise_deep_search():
    B, rounds, sims = 2, 6, 120
    noise = np.random.RandomState(9).dirichlet([10.83 / 116] * 256, size=(B, 4)).astype(np.float32)
    nn = FakeNN(5 * 11 * 11, 2420, seed=32, sharp=2.0)
    orc, eng = _hnef_pair(B, "mt19937", noise=noise, temps=TAFL_TEMPS, add_root_temp=True, add_root_noise=True,
                          max_sims_per_move=sims)
    assert_traces_equal(run_trace(orc, nn, rounds, sims), run_trace(eng, nn, rounds, sims), "hnefatafl noise")
    assert np.array_equal(orc.boards(), eng.boards())
