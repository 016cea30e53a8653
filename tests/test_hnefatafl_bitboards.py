"""The engine's 128-bit hnefatafl rules (alphazero-general_b200/csrc/azb_hnefatafl.cuh), compiled for the HOST by
oracle/t128_host.cpp, against the C oracle (itself pinned to the compiled reference in test_hnefatafl_rules.py): cells,
valid-move masks in ascending action order, win state and observation after every move of random playouts, constructed
captures across the 64-bit seam of the board, and all 8 symmetries with the 2420-entry policy permutation.  No GPU: the
header is plain scalar code, so what a kernel will execute is checked here before the engine serves 121-cell boards."""
import ctypes as C
import os

import numpy as np
import pytest

import _orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HN = _orc.GAME_HNEFATAFL
A, CELLS = 2420, 121


@pytest.fixture(scope="module")
def t128():
    path = os.path.join(ROOT, "oracle", "libt128.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libt128.so"])
    L = C.CDLL(path)
    L.t128_rules_play_from.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    L.t128_symmetry.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bits(L, cells0, turns, acts):
    acts = np.ascontiguousarray(acts, dtype=np.int32)
    cells = np.zeros(CELLS, np.int8); valid = np.zeros(A, np.uint8); win = np.zeros(3, np.uint8)
    obs = np.zeros((5, 11, 11), np.float32); flags = C.c_int32()
    c0 = None if cells0 is None else np.ascontiguousarray(cells0, dtype=np.int8)
    L.t128_rules_play_from(None if c0 is None else _p(c0), int(turns), _p(acts), len(acts), _p(cells), _p(valid), _p(win),
                           _p(obs), C.byref(flags))
    return cells, valid, win, obs, flags.value


def test_action_codec_round_trips(t128):
    assert t128.t128_codec_roundtrip() == 0


def test_random_playouts_equal_the_oracle(t128):
    rs = np.random.RandomState(7)
    outcomes, plies = set(), 0
    for game in range(25):
        acts = []
        while True:
            rc, cells, valid, win, obs = _orc.rules_play(HN, acts)
            b = _bits(t128, None, 0, acts)
            assert rc == 0
            assert np.array_equal(b[0], cells), acts
            assert np.array_equal(b[1], valid), acts          # also: 255 in [0] would flag a non-ascending candidate order
            assert np.array_equal(b[2], win), acts
            assert np.array_equal(b[3], obs), acts
            if win.any():
                outcomes.add(int(np.argmax(win)))
                break
            acts.append(int(rs.choice(np.nonzero(valid)[0])))
        plies += len(acts)
    assert {1, 2} <= outcomes and plies > 2000


def _board(rows):
    return np.array([[int(c) for c in r] for r in rows], dtype=np.int8).ravel()


def _both(t128, cells0, turns, acts):
    o = _orc.rules_from_cells(HN, cells0, turns, acts)
    b = _bits(t128, cells0, turns, acts)
    assert o[0] == 0
    for k in range(4):
        assert np.array_equal(b[k], o[k + 1]), k
    return o, b


def _act(x, y, nx, ny):
    mt = (ny if ny < y else ny - 1) if x == nx else (10 + nx if nx < x else 9 + nx)
    return 20 * (x + 11 * y) + mt


def test_constructed_positions_across_the_word_seam(t128):
    # bit 64 is square (9, 5): rows 5 and 6 straddle the lo / hi words.  Side 2 to move (turns even).
    rows = ["50000000005", "00000000000", "00000000000", "00000000000", "00000000020", "00000400210",
            "00000000020", "00000000000", "00000300000", "00000000000", "50000000005"]
    # (9,5) is a side-1 piece with side 2 on three sides: (10,5) is free.  Moving 2 to (10,5) closes the group -> captured
    cells0 = _board(rows)
    cells0[4 * 11 + 10] = 2                                    # a side-2 piece at (10, 4) that steps down to (10, 5)
    o, b = _both(t128, cells0, 0, [])
    o, b = _both(t128, cells0, 0, [_act(10, 4, 10, 5)])
    assert o[1][5 * 11 + 9] == 0                               # sandwich (8,5)-(10,5) AND surround: the piece is gone
    # four-sided king capture next to the seam, and no two-sided sandwich of the king
    rows = ["50000000005", "00000000000", "00000000000", "00000000000", "00000000020", "00000400230",
            "00000000000", "00000000020", "00000000000", "10000000000", "50000000005"]
    cells0 = _board(rows)
    o, b = _both(t128, cells0, 0, [])
    assert not o[3].any()
    o, b = _both(t128, cells0, 0, [_act(9, 7, 9, 6)])          # third side: (10,5) still open -> no capture
    assert not o[3].any()
    cells0[6 * 11 + 9] = 2; cells0[7 * 11 + 9] = 0; cells0[3 * 11 + 10] = 2
    o, b = _both(t128, cells0, 0, [_act(10, 3, 10, 4)])        # not adjacent yet
    assert not o[3].any()
    o, b = _both(t128, cells0, 2, [_act(10, 3, 10, 4), _act(0, 9, 1, 9), _act(10, 4, 10, 5)])
    assert o[3].tolist() == [1, 0, 0]                          # all four neighbours side 2: side 2 (env player 0) wins
    # the king escapes to a corner; a plain piece may not enter it; an empty throne is crossed but not entered
    rows = ["50000000035", "00000000000", "00000000000", "00000000000", "00000000000", "00100400000",
            "00000000000", "00000000000", "00000000000", "20000000000", "50000000002"]
    cells0 = _board(rows)
    o, b = _both(t128, cells0, 1, [])                          # side 1 to move
    assert o[2][_act(9, 0, 10, 0)] == 1                        # king -> corner
    assert o[2][_act(2, 5, 9, 5)] == 1 and o[2][_act(2, 5, 5, 5)] == 0      # over the empty throne, not onto it
    o, b = _both(t128, cells0, 1, [_act(9, 0, 10, 0)])
    assert o[3].tolist() == [0, 1, 0] and o[1][10] == 8
    o, b = _both(t128, cells0, 0, [])                          # side 2 to move: (10,10) piece next to nothing special
    assert o[2][_act(0, 9, 0, 10)] == 0                        # a plain piece may not enter the corner (0, 10)


def test_symmetries_equal_the_oracle(t128):
    rs = np.random.RandomState(3)
    acts = []
    for _ in range(40):
        rc, cells, valid, win, obs = _orc.rules_play(HN, acts)
        if win.any():
            break
        acts.append(int(rs.choice(np.nonzero(valid)[0])))
    rc, cells, valid, win, obs = _orc.rules_play(HN, acts)
    pi = rs.rand(A).astype(np.float32) * valid
    L = _orc.lib()
    for k in range(8):
        c2 = np.zeros(CELLS, np.int8); p2 = np.zeros(A, np.float32)
        t128.t128_symmetry(_p(cells), len(acts), _p(pi), k, _p(c2), _p(p2))
        oc, op = _orc.symmetry(HN, cells, len(acts), pi, k)
        assert np.array_equal(c2, oc) and np.array_equal(p2, op), k


def test_policy_sum_plan_equals_numpy(t128):
    """np.sum over float32[2420] as a warp will evaluate it (32 leaves of NumPy's pairwise recursion, one per lane, then
    five xor-shuffle steps) is bit-equal to NumPy -- and to the oracle's restatement of the recursion."""
    t128.t128_np_sum_2420.restype = C.c_float
    t128.t128_np_sum_2420.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    off, ln = np.zeros(32, np.int32), np.zeros(32, np.int32)
    rs = np.random.RandomState(11)
    L = _orc.lib()
    L.orc_np_sum_f32.restype = C.c_float
    L.orc_np_sum_f32.argtypes = [C.c_void_p, C.c_int]
    for trial in range(200):
        a = (rs.rand(A) ** (1 + trial % 5)).astype(np.float32)
        if trial % 3 == 0:
            a *= (rs.rand(A) < 0.05)                          # a masked policy: mostly zeros, as after pi *= valids
        got = t128.t128_np_sum_2420(_p(a), _p(off), _p(ln))
        assert np.float32(got) == np.sum(a, dtype=np.float32), trial
        assert np.float32(got) == np.float32(L.orc_np_sum_f32(_p(a), A)), trial
    assert off[0] == 0 and np.array_equal(off[1:], np.cumsum(ln)[:-1]) and int(ln.sum()) == A
    assert set(ln.tolist()) <= {72, 80, 84} and all(int(v) % 8 == 0 for v in off)
