"""ctypes binding of the CPU oracle (oracle/liborc.so).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liborc.so")

GAME_CONNECT4, GAME_BRANDUBH, GAME_HNEFATAFL = 0, 1, 2
RNG_MT19937, RNG_PHILOX = 0, 1


class OrcArgs(C.Structure):
    _fields_ = [
        ("game", C.c_int32), ("num_slots", C.c_int32), ("rng_mode", C.c_int32),
        ("add_root_noise", C.c_int32), ("add_root_temp", C.c_int32),
        ("symmetric_samples", C.c_int32), ("mcts_reset_threshold", C.c_int32),
        ("temp_table_len", C.c_int32), ("games_per_iteration", C.c_int64),
        ("game_id_base", C.c_int64), ("seed", C.c_uint64),
        ("cpuct", C.c_float), ("fpu_reduction", C.c_float),
        ("root_noise_frac", C.c_float), ("root_policy_temp", C.c_float),
        ("temp_table", C.POINTER(C.c_double)), ("mt_seeds", C.POINTER(C.c_uint32)),
        ("arena", C.c_int32),
    ]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "sims", "sum_depth", "sum_children", "nodes_created", "terminal_leaves",
        "games_played", "results", "samples", "moves")]


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcArgs)]
        for name in ("orc_destroy", "orc_generate_batch", "orc_process_batch", "orc_play_moves",
                     "orc_set_root_noise", "orc_root_counts", "orc_last_actions", "orc_turns",
                     "orc_boards", "orc_players", "orc_get_stats", "orc_get_samples", "orc_clear_samples",
                     "orc_get_results"):
            getattr(L, name).restype = None
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_action_size.argtypes = [C.c_void_p]
        L.orc_obs_size.argtypes = [C.c_void_p]
        L.orc_generate_batch.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_play_moves.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_root_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_root_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_last_actions.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_turns.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_boards.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_players.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_stats.argtypes = [C.c_void_p, C.POINTER(OrcStats)]
        L.orc_num_samples.restype = C.c_int64
        L.orc_num_samples.argtypes = [C.c_void_p]
        L.orc_get_samples.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_clear_samples.argtypes = [C.c_void_p]
        L.orc_num_results.restype = C.c_int64
        L.orc_num_results.argtypes = [C.c_void_p]
        L.orc_get_results.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.orc_rules_play.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_rules_play_from.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.orc_c4_win_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_np_sum_f32.restype = C.c_float
        L.orc_np_sum_f32.argtypes = [C.c_void_p, C.c_int]
        L.orc_pow_f32.restype = C.c_float
        L.orc_pow_f32.argtypes = [C.c_float, C.c_float]
        L.orc_mt_seed.argtypes = [C.c_uint32, C.c_void_p]
        L.orc_mt_next.restype = C.c_uint32
        L.orc_mt_next.argtypes = [C.c_void_p]
        L.orc_philox_words.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_mt_shuffle.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_mt_double.restype = C.c_double
        L.orc_mt_double.argtypes = [C.c_void_p]
    return _lib


def rules_from_cells(game, cells0, turns, actions):
    """Like rules_play, from a constructed position (tafl cell codes, `turns` plies played)."""
    d = GAME_DIMS[game]
    acts = np.ascontiguousarray(actions, dtype=np.int32)
    c0 = np.ascontiguousarray(cells0, dtype=np.int8)
    assert c0.size == d["cells"]
    cells = np.zeros(d["cells"], dtype=np.int8)
    valid = np.zeros(d["A"], dtype=np.uint8)
    win = np.zeros(3, dtype=np.uint8)
    obs = np.zeros(d["obs"], dtype=np.float32)
    rc = lib().orc_rules_play_from(game, _p(c0), int(turns), _p(acts), len(acts), _p(cells), _p(valid), _p(win), _p(obs))
    return rc, cells, valid, win, obs


def symmetry(game, cells0, turns, pi, k):
    """Game.symmetries entry k of (position given as cell codes, pi) -> (cells, pi)."""
    d = GAME_DIMS[game]
    c0 = np.ascontiguousarray(cells0, dtype=np.int8)
    p0 = np.ascontiguousarray(pi, dtype=np.float32)
    cells = np.zeros(d["cells"], dtype=np.int8)
    p2 = np.zeros(d["A"], dtype=np.float32)
    L = lib()
    L.orc_rules_symmetry.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert L.orc_rules_symmetry(game, _p(c0), int(turns), _p(p0), int(k), _p(cells), _p(p2)) == 0
    return cells, p2


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


GAME_DIMS = {GAME_CONNECT4: dict(A=7, obs=(4, 6, 7), cells=42),
             GAME_BRANDUBH: dict(A=588, obs=(5, 7, 7), cells=49),
             GAME_HNEFATAFL: dict(A=2420, obs=(5, 11, 11), cells=121)}


def temp_table(temp_scaling_fn, start_temp, max_turns, n=512):
    """Temperature used for a move made at turn t: SelfPlayAgent.playMoves
    iterates temps[i] = fn(temps[i], turns, max_turns) once per move
    (SelfPlayAgent.pyx:156-158), so it is a function of t alone."""
    out, cur = [], start_temp
    for t in range(n):
        cur = temp_scaling_fn(cur, t, max_turns)
        out.append(float(cur))
    return np.asarray(out, dtype=np.float64)


def default_temp_scaling(cur_temp, turns, const_max_turns):
    """alphazero/utils.py:19-27 (scale_temp(0.15, 0.2, ...))."""
    if const_max_turns and (turns + 1) % int(0.15 * const_max_turns) == 0:
        return max(0.2, cur_temp / 2)
    return cur_temp


class OracleAgent:
    """Lock-step self-play agent over the C oracle (SelfPlayAgent surface:
    generateBatch / processBatch / playMoves)."""

    def __init__(self, game=GAME_CONNECT4, num_slots=1, rng_mode=RNG_MT19937, seed=0, mt_seeds=None,
                 game_id_base=0, cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1,
                 add_root_noise=False, add_root_temp=False, symmetric_samples=True,
                 mcts_reset_threshold=0, games_per_iteration=1 << 40, temps=None, arena=False, arena_temp=None,
                 player_to_index=None):
        L = lib()
        self.L = L
        self.B = num_slots
        a = OrcArgs()
        a.game, a.num_slots, a.rng_mode = game, num_slots, rng_mode
        a.add_root_noise, a.add_root_temp = int(add_root_noise), int(add_root_temp)
        a.symmetric_samples, a.mcts_reset_threshold = int(symmetric_samples), int(mcts_reset_threshold or 0)
        a.games_per_iteration, a.game_id_base, a.seed = games_per_iteration, game_id_base, seed
        a.cpuct, a.fpu_reduction = cpuct, fpu_reduction
        a.root_noise_frac, a.root_policy_temp = root_noise_frac, root_policy_temp
        a.arena = int(arena)
        self.player_to_index = list(player_to_index or [0, 1])
        if arena:                      # playMoves uses args.arenaTemp for every move (SelfPlayAgent.pyx:156-158)
            temps = np.full(1, 0.25 if arena_temp is None else arena_temp, dtype=np.float64)
        if temps is None:
            temps = np.ones(1, dtype=np.float64)
        temps = np.ascontiguousarray(temps, dtype=np.float64)
        a.temp_table_len = len(temps)
        a.temp_table = temps.ctypes.data_as(C.POINTER(C.c_double))
        if mt_seeds is not None:
            ms = np.ascontiguousarray(mt_seeds, dtype=np.uint32)
            assert len(ms) == num_slots
            a.mt_seeds = ms.ctypes.data_as(C.POINTER(C.c_uint32))
        self.h = L.orc_create(C.byref(a))
        assert self.h
        self.A = L.orc_action_size(self.h)
        self.obs_shape = GAME_DIMS[game]["obs"]
        self.ncells = GAME_DIMS[game]["cells"]
        self.obs = np.zeros((num_slots,) + self.obs_shape, dtype=np.float32)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_root_noise(self, noise):
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        assert noise.ndim == 3 and noise.shape[0] == self.B
        self.L.orc_set_root_noise(self.h, _p(noise), noise.shape[1], noise.shape[2])

    def generateBatch(self):
        self.L.orc_generate_batch(self.h, _p(self.obs))
        return self.obs

    def processBatch(self, policy, value):
        policy = np.ascontiguousarray(policy, dtype=np.float32)
        value = np.ascontiguousarray(value, dtype=np.float32)
        assert policy.shape == (self.B, self.A) and value.shape == (self.B, 3)
        self.L.orc_process_batch(self.h, _p(policy), _p(value))

    def playMoves(self, fast=False):
        self.L.orc_play_moves(self.h, int(fast))

    def root_counts(self):
        out = np.zeros((self.B, self.A), dtype=np.int32)
        self.L.orc_root_counts(self.h, _p(out))
        return out

    def last_actions(self):
        out = np.zeros(self.B, dtype=np.int32)
        self.L.orc_last_actions(self.h, _p(out))
        return out

    def turns(self):
        out = np.zeros(self.B, dtype=np.int32)
        self.L.orc_turns(self.h, _p(out))
        return out

    def players(self):
        out = np.zeros(self.B, dtype=np.int32)
        self.L.orc_players(self.h, _p(out))
        return out

    def models(self):
        """arena: index of the model that evaluates each slot's leaf (SelfPlayAgent.pyx:117: player_to_index[player])."""
        return np.asarray(self.player_to_index, dtype=np.int32)[self.players()]

    def boards(self):
        out = np.zeros((self.B, self.ncells), dtype=np.int8)
        self.L.orc_boards(self.h, _p(out))
        return out

    def stats(self):
        st = OrcStats()
        self.L.orc_get_stats(self.h, C.byref(st))
        return {n: getattr(st, n) for n, _ in OrcStats._fields_}

    def samples(self):
        n = self.L.orc_num_samples(self.h)
        obs = np.zeros((n,) + self.obs_shape, dtype=np.float32)
        pi = np.zeros((n, self.A), dtype=np.float32)
        z = np.zeros((n, 3), dtype=np.float32)
        slot = np.zeros(n, dtype=np.int32)
        if n:
            self.L.orc_get_samples(self.h, _p(obs), _p(pi), _p(z), _p(slot))
        return obs, pi, z, slot

    def results(self):
        n = self.L.orc_num_results(self.h)
        slot = np.zeros(n, dtype=np.int32)
        turns = np.zeros(n, dtype=np.int32)
        win = np.zeros((n, 3), dtype=np.uint8)
        if n:
            self.L.orc_get_results(self.h, _p(slot), _p(turns), _p(win))
        return slot, turns, win


def rules_play(game, actions):
    d = GAME_DIMS[game]
    acts = np.ascontiguousarray(actions, dtype=np.int32)
    cells = np.zeros(d["cells"], dtype=np.int8)
    valid = np.zeros(d["A"], dtype=np.uint8)
    win = np.zeros(3, dtype=np.uint8)
    obs = np.zeros(d["obs"], dtype=np.float32)
    rc = lib().orc_rules_play(game, _p(acts), len(acts), _p(cells), _p(valid), _p(win), _p(obs))
    return rc, cells, valid, win, obs
