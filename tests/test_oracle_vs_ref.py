"""The C oracle against the compiled reference itself (oracle/_ref), live.
Skipped where oracle/_ref has not been built (it needs /root/reference)."""
import numpy as np
import pytest

import _orc
import _refdriver
from _fakenn import ArenaNN, FakeNN
from _lockstep import assert_queues_equal, assert_traces_equal, run_trace

pytestmark = pytest.mark.skipif(not _refdriver.available(), reason="oracle/_ref not built")
TEMPS = _orc.temp_table(_orc.default_temp_scaling, 1, 42)


@pytest.mark.parametrize("mode,root_temp", [("warmup", False), ("nn", False), ("nn", True)])
def test_connect4_oracle_equals_reference(mode, root_temp):
    B, seeds = 3, [5, 6, 7]
    nn = FakeNN(4 * 6 * 7, 7, seed=42) if mode == "nn" else None
    # root_temp=False runs the UNMODIFIED reference; True needs the deterministic-pow patch
    ref = _refdriver.RefAgent("connect4", B, mt_seeds=seeds, add_root_temp=root_temp, det_pow=root_temp)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, mt_seeds=seeds, add_root_temp=root_temp, temps=TEMPS)
    assert_traces_equal(run_trace(ref, nn, 48, 16, keep_obs=True), run_trace(orc, nn, 48, 16, keep_obs=True), mode)
    assert_queues_equal(ref, orc, mode)
    assert len(ref.results()[0]) > 0


def test_connect4_oracle_equals_reference_with_fed_noise_and_reset():
    B, seeds = 2, [8, 9]
    noise = np.random.RandomState(3).dirichlet([10.83 / 7] * 7, size=(B, 40)).astype(np.float32)
    nn = FakeNN(4 * 6 * 7, 7, seed=43)
    ref = _refdriver.RefAgent("connect4", B, mt_seeds=seeds, add_root_temp=True, add_root_noise=True,
                              det_pow=True, noise=noise, mcts_reset_threshold=4)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, mt_seeds=seeds, add_root_temp=True, add_root_noise=True,
                           temps=TEMPS, mcts_reset_threshold=4)
    orc.set_root_noise(noise)
    assert_traces_equal(run_trace(ref, nn, 40, 12), run_trace(orc, nn, 40, 12), "noise+reset")
    assert_queues_equal(ref, orc, "noise+reset")


@pytest.mark.parametrize("mode,root_temp", [("warmup", False), ("nn", True)])
def test_brandubh_oracle_equals_reference(mode, root_temp):
    B, seeds = 2, [3, 4]
    temps = _orc.temp_table(_orc.default_temp_scaling, 1, None)
    nn = FakeNN(5 * 7 * 7, 588, seed=5, sharp=1.0) if mode == "nn" else None
    ref = _refdriver.RefAgent("brandubh", B, mt_seeds=seeds, add_root_temp=root_temp, det_pow=root_temp)
    orc = _orc.OracleAgent(_orc.GAME_BRANDUBH, B, mt_seeds=seeds, add_root_temp=root_temp, temps=temps)
    assert_traces_equal(run_trace(ref, nn, 70, 10, keep_obs=True), run_trace(orc, nn, 70, 10, keep_obs=True), mode)
    assert_queues_equal(ref, orc, mode)


@pytest.mark.parametrize("p2i,reset", [([0, 1], None), ([1, 0], 5)])
def test_connect4_arena_oracle_equals_reference(p2i, reset):
    """SelfPlayAgent(_is_arena=True): one tree per player, arenaTemp, both trees re-rooted after every move
    (the idle tree expands -- and shuffles -- its root first when it never visited the played child), results only."""
    B, seeds = 3, [11, 12, 13]
    nets = [FakeNN(4 * 6 * 7, 7, seed=50), FakeNN(4 * 6 * 7, 7, seed=51)]
    ref = _refdriver.RefAgent("connect4", B, mt_seeds=seeds, det_pow=True, arena=True, arena_temp=0.25,
                              player_to_index=p2i, mcts_reset_threshold=reset, games_per_iteration=7)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, mt_seeds=seeds, arena=True, arena_temp=0.25, player_to_index=p2i,
                           mcts_reset_threshold=reset or 0, games_per_iteration=7)
    ta = run_trace(ref, ArenaNN(ref, nets), 200, 9, keep_obs=True, until_games=7)      # the worker loop's exit condition
    tb = run_trace(orc, ArenaNN(orc, nets), 200, 9, keep_obs=True, until_games=7)
    assert_traces_equal(ta, tb, "arena")
    assert_queues_equal(ref, orc, "arena")
    assert len(ref.results()[0]) >= 7 and len(ref.samples()[0]) == 0
    assert ref.stats()["games_played"] == orc.stats()["games_played"] == 7


def test_brandubh_arena_oracle_equals_reference():
    B, seeds = 2, [21, 22]
    nets = [FakeNN(5 * 7 * 7, 588, seed=60, sharp=1.0), FakeNN(5 * 7 * 7, 588, seed=61, sharp=1.0)]
    ref = _refdriver.RefAgent("brandubh", B, mt_seeds=seeds, det_pow=True, arena=True, arena_temp=0.5)
    orc = _orc.OracleAgent(_orc.GAME_BRANDUBH, B, mt_seeds=seeds, arena=True, arena_temp=0.5)
    ta = run_trace(ref, ArenaNN(ref, nets), 40, 6, keep_obs=True)
    tb = run_trace(orc, ArenaNN(orc, nets), 40, 6, keep_obs=True)
    assert_traces_equal(ta, tb, "arena-brandubh")
    assert_queues_equal(ref, orc, "arena-brandubh")
