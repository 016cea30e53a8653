"""CPU tests of the oracle's building blocks against their sources of truth:
the reference's own Connect4 known-answer vectors
(alphazero/envs/connect4/test_connect4.py, restated for the current Board API),
NumPy's legacy RandomState, NumPy's float32 reductions and the Random123
Philox4x32-10 known-answer vectors."""
import ctypes as C
import math

import numpy as np
import pytest

import _orc

L = _orc.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- Connect4 rules: reference test vectors -----------------------------------
def test_simple_moves():                      # test_connect4.py:31-40
    rc, cells, valid, win, obs = _orc.rules_play(_orc.GAME_CONNECT4, [4, 5, 4, 3, 0, 6])
    expected = np.array([[0, 0, 0, 0, 0, 0, 0]] * 4 + [[0, 0, 0, 0, 1, 0, 0], [1, 0, 0, -1, 1, -1, -1]], dtype=np.int8)
    assert rc == 0 and np.array_equal(cells.reshape(6, 7), expected)
    assert not win.any()


def test_overfull_column():                   # test_connect4.py:43-53 (6-row board)
    assert _orc.rules_play(_orc.GAME_CONNECT4, [4] * 6)[0] == 0
    assert _orc.rules_play(_orc.GAME_CONNECT4, [4] * 7)[0] == -1   # reference raises ValueError


@pytest.mark.parametrize("moves,expected", [   # test_connect4.py:56-68
    ([], [1] * 7),
    ([0, 1, 2, 3, 4, 5, 6], [1] * 7),
    ([0, 1, 2, 3, 4, 5, 6] * 5, [1] * 7),
    ([0, 1, 2, 3, 4, 5, 6] * 6, [0] * 7),
    ([0, 1, 2] * 3 + [3, 4, 5, 6] * 6, [1] * 3 + [0] * 4),
])
def test_get_valid_moves(moves, expected):
    rc, cells, valid, win, obs = _orc.rules_play(_orc.GAME_CONNECT4, moves)
    assert rc == 0 and valid.tolist() == expected


END_STATE_BOARDS = [                           # test_connect4.py:97-156 -> winner (+1 / -1 / 0)
    (np.zeros((5, 7)), 0),
    ([[0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 1, 0, 0, 0],
      [0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0]], 1),
    ([[0, 0, 0, 0, 1, 0, 0], [0, 0, 0, 1, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0],
      [0, 0, 0, 0, 0, 0, 0]], 1),
    ([[0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 1, 0, 0],
      [0, 0, 0, 0, 0, 1, 0]], 1),
    ([[0, 0, 0, -1], [0, 0, -1, 0], [0, -1, 0, 0], [-1, 0, 0, 0]], -1),
    ([[0, 0, 0, 0, 1], [0, 0, 0, 1, 0], [0, 0, 1, 0, 0], [0, 1, 0, 0, 0]], 1),
    ([[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 0, 1, 0]], 1),
    ([[0, 0, 0, 0, 0, 0, 0], [0, 0, 0, -1, 0, 0, 0], [0, 0, 0, -1, 0, 0, 1], [0, 0, 0, 1, 1, -1, -1],
      [0, 0, 0, -1, 1, 1, 1], [0, -1, 0, -1, 1, -1, 1]], 0),
    ([[0, 0, 0, 0, 0, 0, 0], [0, 0, 0, -1, 0, 0, 0], [1, 0, 1, -1, 0, 0, 0], [-1, -1, 1, 1, 0, 0, 0],
      [1, 1, 1, -1, 0, 0, 0], [1, -1, 1, -1, 0, -1, 0]], 1),
    ([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0], [0, 0, 0, -1, 0, 0, 0], [0, 0, 1, 1, -1, 0, -1],
      [0, 0, -1, 1, 1, 1, 1], [-1, 0, -1, 1, -1, -1, -1]], 1),
]


@pytest.mark.parametrize("idx", range(len(END_STATE_BOARDS)))
def test_game_ended(idx):
    board, winner = END_STATE_BOARDS[idx]
    b = np.ascontiguousarray(board, dtype=np.int32)
    assert L.orc_c4_win_state(_p(b), b.shape[0], b.shape[1], 4) == winner


def test_draw_and_symmetry_through_samples():
    # mirror symmetry (test_connect4.py:71-94): emitted sample pairs are (state, pi), (mirror, pi[::-1])
    ag = _orc.OracleAgent(_orc.GAME_CONNECT4, 2, mt_seeds=[1, 2], temps=_orc.temp_table(_orc.default_temp_scaling, 1, 42))
    from _fakenn import warmup_outputs
    for _ in range(60):
        for _ in range(8):
            ag.generateBatch()
            ag.processBatch(*warmup_outputs(2, 7))
        ag.playMoves()
    obs, pi, z, slot = ag.samples()
    assert len(obs) >= 4 and len(obs) % 2 == 0
    assert np.array_equal(obs[1::2], obs[0::2][:, :, :, ::-1])
    assert np.array_equal(pi[1::2], pi[0::2][:, ::-1])
    assert np.array_equal(z[1::2], z[0::2]) and np.all(z.sum(1) == 1)
    assert np.allclose(pi.sum(1), 1, atol=1e-5)


# ---- RNG ---------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 123, 2 ** 32 - 1])
def test_mt19937_matches_numpy_legacy(seed):
    st = np.zeros(625, dtype=np.uint32)
    L.orc_mt_seed(seed, _p(st))
    rs = np.random.RandomState(seed)
    ours = [L.orc_mt_next(_p(st)) for _ in range(1500)]
    theirs = rs.randint(0, 2 ** 32, size=1500, dtype=np.uint64).astype(np.uint64)   # one 32-bit draw each
    assert ours == [int(x) for x in theirs]
    for _ in range(50):
        assert L.orc_mt_double(_p(st)) == rs.random_sample()


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 33, 60, 96])
def test_list_shuffle_matches_numpy_legacy(n):
    st = np.zeros(625, dtype=np.uint32)
    L.orc_mt_seed(99 + n, _p(st))
    np.random.seed(99 + n)
    for _ in range(20):
        perm = np.zeros(n, dtype=np.int32)
        L.orc_mt_shuffle(_p(st), n, _p(perm))
        lst = list(range(n))
        np.random.shuffle(lst)               # the untyped (list) path Node.add_children uses
        assert perm.tolist() == lst


def _philox_block(ctr, key):
    """Independent Philox4x32-10 (Salmon et al., Random123) in pure Python."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
        k = [(k[0] + W0) & 0xffffffff, (k[1] + W1) & 0xffffffff]
    return c


def test_philox_known_answers_and_stream_mapping():
    # Random123 kat_vectors for philox4x32-10
    assert _philox_block([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox_block([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox_block([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    out = np.zeros(4, dtype=np.uint32)
    L.orc_philox_words(0, 0, 0, 4, _p(out))
    assert out.tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    rs = np.random.RandomState(4)
    for _ in range(20):
        seed, gid, w = (int(rs.randint(0, 2 ** 63)) for _ in range(3))
        w >>= 4
        n = 9
        out = np.zeros(n, dtype=np.uint32)
        L.orc_philox_words(seed, gid, w, n, _p(out))
        for i in range(n):
            ww = w + i
            blk = _philox_block([(ww >> 2) & 0xffffffff, (ww >> 34) & 0xffffffff, gid & 0xffffffff, gid >> 32],
                                [seed & 0xffffffff, seed >> 32])
            assert int(out[i]) == blk[ww & 3]


# ---- NumPy float32 semantics ----------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 15, 16, 17, 40, 127, 128, 129, 257, 588, 1000])
def test_pairwise_sum_matches_numpy(n):
    rs = np.random.RandomState(n)
    for t in range(50):
        a = rs.random_sample(n).astype(np.float32)
        if t % 2:
            a[rs.random_sample(n) < 0.6] = 0
        assert np.float32(L.orc_np_sum_f32(_p(a), n)) == np.sum(a)


def test_pow_is_correctly_rounded_double_pow():
    rs = np.random.RandomState(0)
    for e in (1.0 / np.float32(1.1), 4.0, 5.0, 2.0, 0.5, 1.0):
        e32 = float(np.float32(e))
        for x in rs.random_sample(200).astype(np.float32):
            want = np.float32(math.pow(float(x), e32)) if e32 != 1.0 else x
            assert np.float32(L.orc_pow_f32(float(x), e32)) == want
    assert L.orc_pow_f32(0.0, 0.9) == 0.0
