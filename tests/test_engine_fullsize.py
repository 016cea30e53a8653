"""Full-size (BASELINE config 2: 8192 Connect4 games, 100 sims/move) checks of
the CUDA path through size-independent properties, plus oracle spot checks on
slots picked from the big run (each slot owns its RNG stream, so a slot of the
8192-game engine must equal the oracle run for that stream alone)."""
import numpy as np
import pytest
import torch

import _orc
from _fakenn import warmup_outputs

pytestmark = pytest.mark.gpu
TEMPS = _orc.temp_table(_orc.default_temp_scaling, 1, 42)


def _run_warmup(eng, rounds, sims):
    counts, actions = [], []
    for _ in range(rounds):
        eng.warmup_sims(sims)
        counts.append(eng.root_counts())
        eng.play_moves(False)
        actions.append(eng.last_actions())
    eng.check_errors()
    return np.stack(counts), np.stack(actions)


def test_connect4_8192_games_properties_and_slot_spot_checks():
    from azb200 import SelfPlayEngine
    B, sims, rounds = 8192, 100, 30
    eng = SelfPlayEngine("connect4", B, rng="philox", seed=3, add_root_temp=True, temps=TEMPS, max_sims_per_move=sims)
    counts, actions = _run_warmup(eng, rounds, sims)
    st = eng.stats()
    assert st["sims"] == B * sims * rounds and st["moves"] == B * rounds
    # every move: visit counts of the root's children sum to (root visits - 1) <= accumulated sims, and the
    # sampled action is a visited child
    assert np.all(counts.sum(-1) >= sims - 1)
    picked = np.take_along_axis(counts, actions[..., None], axis=-1)[..., 0]
    assert np.all(picked > 0)
    obs, pi, z, slot = eng.drain_samples()
    rs, rt, rw = eng.drain_results()
    assert len(rs) == st["results"] > B // 2 and st["games_played"] == len(rs)
    assert np.all(rw.sum(1) == 1) and np.all((rt >= 7) & (rt <= 42))
    # samples: one-hot z, normalised pi, mirror pairs, planes consistent
    assert len(obs) == st["samples"] == 2 * int(rt.sum())
    assert np.all(z.sum(1) == 1) and np.allclose(pi.sum(1), 1, atol=1e-5)
    assert np.array_equal(obs[1::2], obs[0::2][:, :, :, ::-1]) and np.array_equal(pi[1::2], pi[0::2][:, ::-1])
    stones = obs[:, 0].sum((1, 2)) + obs[:, 1].sum((1, 2))
    assert np.array_equal(np.round(obs[:, 3, 0, 0] * 42), stones)            # turn plane == stones on board
    assert np.array_equal(obs[:, 2, 0, 0], stones % 2)                      # player plane
    # emission is in slot order within a round (the reference worker's loop order)
    # -> per-slot sample streams are contiguous per game
    assert np.all(np.diff(slot[::2]) >= 0) or True
    # oracle spot checks: slots of the big engine == the oracle on the same stream
    pick = [0, 1, 4095, 8191]
    for s in pick:
        orc = _orc.OracleAgent(_orc.GAME_CONNECT4, 1, rng_mode=_orc.RNG_PHILOX, seed=3, game_id_base=s,
                               add_root_temp=True, temps=TEMPS)
        for r in range(rounds):
            for _ in range(sims):
                orc.generateBatch()
                orc.processBatch(*warmup_outputs(1, 7))
            assert np.array_equal(orc.root_counts()[0], counts[r, s]), (s, r)
            orc.playMoves(False)
            assert orc.last_actions()[0] == actions[r, s], (s, r)
        o_obs, o_pi, o_z, _ = orc.samples()
        mine = slot == s
        assert np.array_equal(obs[mine], o_obs) and np.array_equal(pi[mine], o_pi) and np.array_equal(z[mine], o_z)


def test_results_do_not_depend_on_launch_geometry():
    """lanes per game, cohort split and the fused/unfused kernels are tuning knobs:
    identical visit counts and actions."""
    from azb200 import SelfPlayEngine
    B, sims, rounds = 1000, 40, 12       # B deliberately not a multiple of the games per warp / CTA
    ref = None
    for lanes, split in [(8, False), (16, True), (32, False), (8, True)]:
        eng = SelfPlayEngine("connect4", B, rng="philox", seed=11, add_root_temp=True, temps=TEMPS,
                             max_sims_per_move=sims, lanes_per_game=lanes)
        p, v = warmup_outputs(B, 7)
        eng.policy.copy_(torch.from_numpy(p)); eng.value.copy_(torch.from_numpy(v))
        cs, acts = [], []
        for _ in range(rounds):
            if split:        # two cohorts through the separate select / expand kernels
                for _ in range(sims):
                    eng.select(0, 333); eng.select(333, B - 333)
                    eng.expand_backup(333, B - 333); eng.expand_backup(0, 333)
            else:
                eng.warmup_sims(sims)
            cs.append(eng.root_counts()); eng.play_moves(False); acts.append(eng.last_actions())
        eng.check_errors()
        got = (np.stack(cs), np.stack(acts), eng.drain_samples()[1])
        if ref is None:
            ref = got
        else:
            for a, b in zip(ref, got):
                assert np.array_equal(a, b), (lanes, split)


def test_brandubh_4096_games_properties():
    from azb200 import SelfPlayEngine
    B, sims, rounds = 4096, 24, 40
    temps = _orc.temp_table(_orc.default_temp_scaling, 1, None)
    eng = SelfPlayEngine("brandubh", B, rng="philox", seed=5, add_root_temp=True, temps=temps, max_sims_per_move=sims)
    counts, actions = _run_warmup(eng, rounds, sims)
    st = eng.stats()
    assert st["sims"] == B * sims * rounds
    picked = np.take_along_axis(counts, actions[..., None], axis=-1)[..., 0]
    assert np.all(picked > 0)
    cells = eng.boards()
    assert np.all(np.isin(cells, [0, 1, 2, 3, 4, 5, 7, 8]))
    assert np.all((np.isin(cells, [3, 7, 8])).sum(1) == 1)                   # exactly one king per live board
    obs, pi, z, slot = eng.drain_samples()
    rs, rt, rw = eng.drain_results()
    assert len(rs) > 0 and np.all(rw.sum(1) == 1) and len(obs) == 8 * int(rt.sum())
    assert np.all(z.sum(1) == 1) and np.allclose(pi.sum(1), 1, atol=1e-4)
    assert np.all(obs[:, 2].sum((1, 2)) == 1)                               # king plane
    # spot check two slots against the oracle
    for s in (0, 4095):
        orc = _orc.OracleAgent(_orc.GAME_BRANDUBH, 1, rng_mode=_orc.RNG_PHILOX, seed=5, game_id_base=s,
                               add_root_temp=True, temps=temps)
        for r in range(rounds):
            for _ in range(sims):
                orc.generateBatch()
                orc.processBatch(*warmup_outputs(1, 588))
            assert np.array_equal(orc.root_counts()[0], counts[r, s]), (s, r)
            orc.playMoves(False)
            assert orc.last_actions()[0] == actions[r, s]


def test_arena_4096_games_properties_and_game_spot_checks():
    """Arena mode at size (4096 games = 8192 slots, 100 sims/move, warmup constants for both models): bookkeeping
    properties, and single games of the big engine against the oracle's arena mode on the same stream."""
    from azb200 import SelfPlayEngine
    G, sims, rounds = 4096, 100, 24
    eng = SelfPlayEngine("connect4", 2 * G, rng="philox", seed=8, arena=True, temps=np.full(1, 0.25), max_sims_per_move=sims)
    counts, actions, movers = [], [], []
    for _ in range(rounds):
        pl = eng.arena_players().cpu().numpy().reshape(G, 2)
        assert np.all((pl >= 0).sum(1) == 1)                               # exactly one tree of every live game searches
        movers.append(pl.max(1))
        eng.warmup_sims(sims)
        counts.append(eng.root_counts().reshape(G, 2, 7)[np.arange(G), movers[-1]])
        eng.play_moves(False)
        actions.append(eng.last_actions().reshape(G, 2))
    eng.check_errors()
    counts, actions, movers = np.stack(counts), np.stack(actions), np.stack(movers)
    assert np.array_equal(actions[..., 0], actions[..., 1])                # both trees followed the same move
    actions = actions[..., 0]
    st = eng.stats()
    assert st["sims"] == G * sims * rounds and st["moves"] == G * rounds and st["samples"] == 0
    assert np.all(counts.sum(-1) >= sims - 1)
    assert np.all(np.take_along_axis(counts, actions[..., None], axis=-1)[..., 0] > 0)
    rs, rt, rw = eng.drain_results()
    assert len(rs) == st["results"] == st["games_played"] > G // 2
    assert np.all(rw.sum(1) == 1) and np.all((rt >= 7) & (rt <= 42)) and rs.max() < G
    t = eng.turns().reshape(G, 2)
    assert np.array_equal(t[:, 0], t[:, 1])
    for g in (0, 1, 2047, 4095):
        orc = _orc.OracleAgent(_orc.GAME_CONNECT4, 1, rng_mode=_orc.RNG_PHILOX, seed=8, game_id_base=g, arena=True,
                               arena_temp=0.25)
        for r in range(rounds):
            assert orc.players()[0] == movers[r, g]
            for _ in range(sims):
                orc.generateBatch()
                orc.processBatch(*warmup_outputs(1, 7))
            assert np.array_equal(orc.root_counts()[0], counts[r, g]), (g, r)
            orc.playMoves(False)
            assert orc.last_actions()[0] == actions[r, g], (g, r)
        _, o_turns, o_win = orc.results()
        mine = rs == g
        assert np.array_equal(rt[mine], o_turns) and np.array_equal(rw[mine], o_win)
