"""Brandubh rules of the C oracle against the compiled reference on random
playouts (the reference ships no tafl test vectors: SURVEY section 4), plus the
documented rule quirks as explicit cases."""
import numpy as np
import pytest

import _orc
import _refdriver

needs_ref = pytest.mark.skipif(not _refdriver.available(), reason="oracle/_ref not built")


def test_start_position():
    rc, cells, valid, win, obs = _orc.rules_play(_orc.GAME_BRANDUBH, [])
    want = ["5002005", "0002000", "0001000", "2217122", "0001000", "0002000", "5002005"]
    assert rc == 0 and cells.reshape(7, 7).tolist() == [[int(c) for c in r] for r in want]
    assert int(valid.sum()) == 40 and not win.any()          # SURVEY: 40 legal moves at the start
    assert obs.shape == (5, 7, 7) and obs[3].max() == 0 and obs[4].max() == 0
    assert obs[0].sum() == 8 and obs[1].sum() == 4 and obs[2].sum() == 1


@needs_ref
def test_random_playouts_match_reference():
    G = _refdriver.game_class("brandubh")
    rs = np.random.RandomState(0)
    outcomes = set()
    for game in range(40):
        g, acts = G(), []
        while True:
            rc, cells, valid, win, obs = _orc.rules_play(_orc.GAME_BRANDUBH, acts)
            assert rc == 0
            rv, rw = g.valid_moves(), g.win_state()
            assert np.array_equal(cells, np.asarray(g._board._state).astype(np.int8).ravel()), acts
            assert np.array_equal(win, rw), acts
            assert np.array_equal(valid, rv), acts
            assert np.array_equal(obs, g.observation()), acts
            if rw.any():
                outcomes.add(int(np.argmax(rw)))
                break
            a = int(rs.choice(np.nonzero(rv)[0]))
            g.play_action(a)
            acts.append(a)
    assert outcomes == {0, 1, 2}


@needs_ref
def test_symmetries_match_reference():
    G = _refdriver.game_class("brandubh")
    rs = np.random.RandomState(1)
    ag = _orc.OracleAgent(_orc.GAME_BRANDUBH, 1, mt_seeds=[1])
    # drive the oracle's sample path: the emitted 8-fold samples must equal Game.symmetries
    from _fakenn import warmup_outputs
    g = G()
    hist = []
    for _ in range(200):
        for _ in range(4):
            ag.generateBatch()
            ag.processBatch(*warmup_outputs(1, 588))
        counts = ag.root_counts()[0].astype(np.float32)
        pi = counts / counts.sum()
        pi = pi / pi.sum()
        hist.append((g.clone(), pi))
        ag.playMoves()
        a = int(ag.last_actions()[0])
        g.play_action(a)
        if g.win_state().any():
            break
    obs, pis, z, slot = ag.samples()
    assert len(obs) == 8 * len(hist)
    i = 0
    for st, pi in hist:
        for s2, pi2 in st.symmetries(pi):
            assert np.array_equal(obs[i], s2.observation())
            assert np.array_equal(pis[i], pi2)
            i += 1
