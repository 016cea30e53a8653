"""Hnefatafl (11x11) in the C oracle against the compiled reference: the rules on random playouts
(cells, valid moves, win state, observation after every move), the 8-fold symmetric samples and the MCTS / self-play
lock-step.  Groundwork for the engine: the reference ships no tafl test vectors, so the oracle is
pinned by differential testing only; the CUDA rules for 121-cell boards are not written yet."""
import numpy as np
import pytest

import _orc
import _refdriver
from _fakenn import FakeNN, warmup_outputs
from _lockstep import assert_queues_equal, assert_traces_equal, run_trace

needs_ref = pytest.mark.skipif(not _refdriver.available(), reason="oracle/_ref not built")
HN = _orc.GAME_HNEFATAFL


def test_start_position():
    rc, cells, valid, win, obs = _orc.rules_play(HN, [])
    want = ["50022222005", "00000200000", "00000000000", "20000100002", "20001110002", "22011711022",
            "20001110002", "20000100002", "00000000000", "00000200000", "50022222005"]          # fastafl/variants.py:1-11
    assert rc == 0 and cells.reshape(11, 11).tolist() == [[int(c) for c in r] for r in want]
    assert valid.shape == (2420,) and int(valid.sum()) == 116 and not win.any()
    assert obs.shape == (5, 11, 11) and obs[0].sum() == 24 and obs[1].sum() == 12 and obs[2].sum() == 1
    assert obs[3].max() == 0 and obs[4].max() == 0            # side 2 to move: 2 - to_play = 0


@needs_ref
def test_random_playouts_match_reference():
    G = _refdriver.game_class("hnefatafl")
    assert G.action_size() == 2420 and tuple(G.observation_size()) == (5, 11, 11)
    rs = np.random.RandomState(0)
    outcomes, plies = set(), 0
    for game in range(12):
        g, acts = G(), []
        while True:
            rc, cells, valid, win, obs = _orc.rules_play(HN, acts)
            assert rc == 0
            rv, rw = g.valid_moves(), g.win_state()
            assert np.array_equal(cells, np.asarray(g._board._state).astype(np.int8).ravel()), acts
            assert np.array_equal(win, rw), acts
            assert np.array_equal(valid, rv), acts
            if len(acts) % 7 == 0 or rw.any():
                assert np.array_equal(obs, g.observation()), acts
            if rw.any():
                outcomes.add(int(np.argmax(rw)))
                break
            a = int(rs.choice(np.nonzero(rv)[0]))
            g.play_action(a)
            acts.append(a)
        plies += len(acts)
    assert {1, 2} <= outcomes and plies > 500                # king's side wins and 512-ply draws; side 2: next test


@needs_ref
def test_four_sided_king_capture_and_no_sandwich():
    """king_two_sided_capture = False (variants.py:21): a king between two pieces of side 2 is NOT taken; the king is
    captured when every in-bounds neighbour is side 2 / throne / escape (Board.king_captured, cengine.pyx:153-161)."""
    from fastafl.cengine import Board
    from fastafl import variants
    G = _refdriver.game_class("hnefatafl")
    rows = ["50000000005", "00000000000", "00000000000", "00000000000", "00002000000", "00023200000",
            "00000000000", "00002000000", "00000000000", "20000000001", "50000000005"]
    # reference: the position as the Board's state, side 2 moves (7,... ) -> the fourth neighbour
    b = Board("\n".join(rows), *variants.hnefatafl_args[1:])
    g = G(b)
    assert not g.win_state().any()                           # three sides: no capture, and no two-sided sandwich
    a = [i for i in np.nonzero(g.valid_moves())[0]]
    # find the move 2: (4,7) -> (4,6) through the reference's own codec
    from alphazero.envs.hnefatafl.fastafl import get_action
    from boardgame import Square
    act = get_action(b, (Square(4, 7), Square(4, 6)))
    assert act in a
    g.play_action(act)
    assert g.win_state().tolist() == [1, 0, 0]               # winner 2 -> index 2 - 2 = 0
    # the oracle from the same cells (through its single-tree entry: set the state, ask for the winner)
    cells = np.array([[int(c) for c in r] for r in rows], dtype=np.int8).ravel()
    w0, w1 = _orc.rules_from_cells(HN, cells, 0, []), _orc.rules_from_cells(HN, cells, 0, [act])
    assert not w0[3].any() and w1[3].tolist() == [1, 0, 0]
    assert np.array_equal(w1[1], np.asarray(g._board._state).astype(np.int8).ravel())


@needs_ref
def test_symmetries_match_reference():
    G = _refdriver.game_class("hnefatafl")
    ag = _orc.OracleAgent(HN, 1, mt_seeds=[1])
    g, hist = G(), []
    for _ in range(12):
        for _ in range(3):
            ag.generateBatch()
            ag.processBatch(*warmup_outputs(1, 2420))
        counts = ag.root_counts()[0].astype(np.float32)
        pi = counts / counts.sum()
        hist.append((g.clone(), pi / pi.sum()))
        ag.playMoves()
        g.play_action(int(ag.last_actions()[0]))
    # samples are emitted when a game ends: finish this one quickly through the reference-equal path
    while not g.win_state().any():
        for _ in range(2):
            ag.generateBatch()
            ag.processBatch(*warmup_outputs(1, 2420))
        counts = ag.root_counts()[0].astype(np.float32)
        pi = counts / counts.sum()
        hist.append((g.clone(), pi / pi.sum()))
        ag.playMoves()
        g.play_action(int(ag.last_actions()[0]))
    obs, pis, z, slot = ag.samples()
    assert len(obs) == 8 * len(hist)
    i = 0
    for n, (st, pi) in enumerate(hist):                       # the first six moves and the last two, all 8 symmetries each
        if n >= 6 and n < len(hist) - 2:
            i += 8
            continue
        for s2, pi2 in st.symmetries(pi):
            assert np.array_equal(obs[i], s2.observation())
            assert np.array_equal(pis[i], pi2)
            i += 1


@needs_ref
@pytest.mark.parametrize("mode,root_temp", [("warmup", False), ("nn", True)])
def test_oracle_equals_reference_in_lock_step(mode, root_temp):
    B, seeds = 2, [3, 4]
    temps = _orc.temp_table(_orc.default_temp_scaling, 1, None)
    nn = FakeNN(5 * 11 * 11, 2420, seed=5, sharp=1.0) if mode == "nn" else None
    ref = _refdriver.RefAgent("hnefatafl", B, mt_seeds=seeds, add_root_temp=root_temp, det_pow=root_temp)
    orc = _orc.OracleAgent(HN, B, mt_seeds=seeds, add_root_temp=root_temp, temps=temps)
    assert_traces_equal(run_trace(ref, nn, 30, 6, keep_obs=True), run_trace(orc, nn, 30, 6, keep_obs=True), mode)
    assert_queues_equal(ref, orc, mode)
