"""SelfPlayEngine -- thin Python owner of one libazb200 engine.

PyTorch is used only as plumbing: device pointers exported by the C ABI are
wrapped as tensors (``__cuda_array_interface__``) so a caller's network can
read the observation batch and write policy / value rows in place, and the
current torch stream is the stream kernels are enqueued on.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import AzbConfig, AzbStats, check

GAMES = {"connect4": _capi.GAME_CONNECT4, "brandubh": _capi.GAME_BRANDUBH, "hnefatafl": _capi.GAME_HNEFATAFL}


def default_temp_scaling(cur_temp, turns, const_max_turns):
    """alphazero/utils.py:19-27: halve every int(0.15*max_turns) plies, floor 0.2."""
    if const_max_turns and (turns + 1) % int(0.15 * const_max_turns) == 0:
        return max(0.2, cur_temp / 2)
    return cur_temp


def temp_table(temp_scaling_fn, start_temp, max_turns, n=512):
    """Temperature of the move made at turn t.  SelfPlayAgent.playMoves applies
    ``temps[i] = fn(temps[i], turns, max_turns)`` once per move starting from
    startTemp (SelfPlayAgent.pyx:156-158,199), so it is a function of t alone."""
    out, cur = [], start_temp
    for t in range(n):
        cur = temp_scaling_fn(cur, t, max_turns)
        out.append(float(cur))
    return np.asarray(out, dtype=np.float64)


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None,
        }


class SelfPlayEngine:
    def __init__(self, game="connect4", num_games=1, device=0, rng="philox", seed=0, mt_seeds=None,
                 game_id_base=0, cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1,
                 add_root_noise=False, add_root_temp=False, symmetric_samples=True,
                 mcts_reset_threshold=None, games_per_iteration=0, max_sims_per_move=100,
                 max_nodes_per_game=0, sample_capacity=0, temps=None, lanes_per_game=0, arena=False):
        self.lib = _capi.load()
        self.h = C.c_void_p()
        cfg = AzbConfig()
        cfg.abi_version = _capi.ABI_VERSION
        cfg.game = GAMES[game] if isinstance(game, str) else int(game)
        cfg.num_games, cfg.device = int(num_games), int(device)
        cfg.rng_mode = {"mt19937": _capi.RNG_MT19937, "philox": _capi.RNG_PHILOX}[rng]
        cfg.add_root_noise, cfg.add_root_temp = int(bool(add_root_noise)), int(bool(add_root_temp))
        cfg.symmetric_samples = int(bool(symmetric_samples))
        cfg.mcts_reset_threshold = int(mcts_reset_threshold or 0)
        cfg.max_sims_per_move, cfg.max_nodes_per_game = int(max_sims_per_move), int(max_nodes_per_game)
        cfg.games_per_iteration, cfg.sample_capacity = int(games_per_iteration or 0), int(sample_capacity)
        cfg.game_id_base, cfg.seed = int(game_id_base), int(seed)
        cfg.lanes_per_game = int(lanes_per_game)
        cfg.arena = int(bool(arena))
        self.arena = bool(arena)
        self.leaf_dedup = False
        cfg.cpuct, cfg.fpu_reduction = float(cpuct), float(fpu_reduction)
        cfg.root_noise_frac, cfg.root_policy_temp = float(root_noise_frac), float(root_policy_temp)
        if temps is not None:
            self._temps = np.ascontiguousarray(temps, dtype=np.float64)
            cfg.temp_table_len = len(self._temps)
            cfg.temp_table = self._temps.ctypes.data_as(C.POINTER(C.c_double))
        if mt_seeds is not None:
            self._mt = np.ascontiguousarray(mt_seeds, dtype=np.uint32)
            if len(self._mt) != num_games:
                raise ValueError("mt_seeds needs one entry per game")
            cfg.mt_seeds = self._mt.ctypes.data_as(C.POINTER(C.c_uint32))
        check(self.lib.azb_create(C.byref(cfg), C.byref(self.h)))
        self.device = int(device)
        self.quota = int(games_per_iteration) if games_per_iteration else (1 << 62)
        self.B = int(num_games)
        self.A = self.lib.azb_action_size(self.h)
        chw = (C.c_int32 * 3)()
        check(self.lib.azb_observation_size(self.h, C.byref(chw)))
        self.obs_shape = tuple(int(x) for x in chw)
        self.ncells = self.obs_shape[1] * self.obs_shape[2]
        self._obs = self._policy = self._value = None

    # ---- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.azb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset_games(self, seed=0, mt_seeds=None):
        p = None
        if mt_seeds is not None:
            self._mt = np.ascontiguousarray(mt_seeds, dtype=np.uint32)
            p = self._mt.ctypes.data_as(C.c_void_p)
        check(self.lib.azb_reset_games(self.h, int(seed), p))

    def set_quota(self, games_per_iteration):
        check(self.lib.azb_set_quota(self.h, int(games_per_iteration or 0)))
        self.quota = int(games_per_iteration) if games_per_iteration else (1 << 62)

    # ---- NN I/O buffers as torch tensors (zero copy) -----------------------
    def _wrap(self, ptr, shape, typestr="<f4"):
        import torch
        return torch.as_tensor(_DevArray(ptr, shape, typestr), device=f"cuda:{self.device}")

    @property
    def nn_rows(self):
        """device int32 [B]: slots whose current leaf needs the network (written by select)."""
        if getattr(self, "_nn_rows", None) is None:
            self._nn_rows = self._wrap(self.lib.azb_nn_rows_ptr(self.h), (self.B,), "<i4")
        return self._nn_rows

    def nn_count_ptr(self):
        """Device address of the int32 count of valid nn_rows entries written by the LAST select call (two counters
        alternate between consecutive selects; query after every select)."""
        return int(self.lib.azb_nn_count_ptr(self.h))

    @property
    def obs(self):
        if self._obs is None:
            self._obs = self._wrap(self.lib.azb_obs_ptr(self.h), (self.B,) + self.obs_shape)
        return self._obs

    @property
    def policy(self):
        if self._policy is None:
            self._policy = self._wrap(self.lib.azb_policy_ptr(self.h), (self.B, self.A))
        return self._policy

    @property
    def value(self):
        if self._value is None:
            self._value = self._wrap(self.lib.azb_value_ptr(self.h), (self.B, 3))
        return self._value

    @staticmethod
    def _stream(stream):
        if stream is None:
            import torch
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if hasattr(stream, "cuda_stream"):
            return C.c_void_p(stream.cuda_stream)
        return C.c_void_p(int(stream))

    # ---- hot path -----------------------------------------------------------
    def select(self, first=0, count=0, stream=None):
        check(self.lib.azb_select(self.h, first, count, self._stream(stream)))

    def expand_backup(self, first=0, count=0, policy=None, value=None, stream=None):
        pp = C.c_void_p(policy.data_ptr()) if policy is not None else None
        vp = C.c_void_p(value.data_ptr()) if value is not None else None
        if policy is not None:
            assert policy.is_cuda and policy.is_contiguous() and policy.dtype.is_floating_point and policy.element_size() == 4
            assert tuple(policy.shape) == (self.B, self.A)
        if value is not None:
            assert value.is_cuda and value.is_contiguous() and value.element_size() == 4
            assert tuple(value.shape) == (self.B, 3)
        check(self.lib.azb_expand_backup(self.h, first, count, pp, vp, self._stream(stream)))

    def expand_backup_select(self, first=0, count=0, stream=None):
        """expand_backup of this simulation + select of the next, one launch (engine-owned policy / value rows)."""
        check(self.lib.azb_expand_backup_select(self.h, first, count, None, None, self._stream(stream)))

    def play_moves(self, fast=False, stream=None):
        check(self.lib.azb_play_moves(self.h, int(bool(fast)), self._stream(stream)))

    # ---- single-tree API (azb200.mcts.MCTS) --------------------------------------------------------------
    def set_state(self, slot, cells, turns, stream=None):
        cells = np.ascontiguousarray(cells, dtype=np.int8).ravel()
        assert cells.size == self.ncells
        check(self.lib.azb_set_state(self.h, int(slot), cells.ctypes.data_as(C.c_void_p), int(turns), self._stream(stream)))

    def force_move(self, slot, action, stream=None):
        check(self.lib.azb_force_move(self.h, int(slot), int(action), self._stream(stream)))

    def set_root_flags(self, add_root_noise, add_root_temp):
        check(self.lib.azb_set_root_flags(self.h, int(bool(add_root_noise)), int(bool(add_root_temp))))

    def set_leaf_dedup(self, on=True):
        """Games whose leaves have the same observation share one network evaluation (include/azb200.h:
        azb_set_leaf_dedup): select lists only the first of them in nn_rows, expand/backup reads its rows for the others.
        Bit-identical results; engine-owned policy / value rows and compact evaluation only."""
        check(self.lib.azb_set_leaf_dedup(self.h, int(bool(on))))
        self.leaf_dedup = bool(on)

    def duplicate_leaves(self):
        """leaves served by another game's evaluation since creation / reset"""
        out = C.c_int64(0)
        check(self.lib.azb_duplicate_leaves(self.h, C.byref(out)))
        return int(out.value)

    def arena_set_player_to_index(self, player_to_index):
        """arena: SelfPlayAgent.player_to_index = [m, 1 - m]; the per-model row lists follow it."""
        check(self.lib.azb_arena_set_player_to_index(self.h, int(player_to_index[0])))

    def arena_rows(self, model):
        """arena: device int32 [B / 2] view of the rows model `model` has to evaluate (after select)."""
        return self._wrap(self.lib.azb_arena_rows_ptr(self.h, int(model)), (self.B // 2,), "<i4")

    def arena_count_ptr(self, model):
        return int(self.lib.azb_arena_count_ptr(self.h, int(model)))

    def arena_players(self, stream=None):
        """arena mode: device int32 [B]: env player whose tree searches in each slot this round, -1 = idle slot."""
        import torch
        if getattr(self, "_arena_players", None) is None:
            self._arena_players = torch.empty(self.B, dtype=torch.int32, device=f"cuda:{self.device}")
        check(self.lib.azb_arena_players(self.h, C.c_void_p(self._arena_players.data_ptr()), self._stream(stream)))
        return self._arena_players

    def warmup_sims(self, sims, stream=None):
        check(self.lib.azb_warmup_sims(self.h, int(sims), self._stream(stream)))

    def set_root_noise(self, noise):
        if noise is None:
            check(self.lib.azb_set_root_noise(self.h, None, 0, 0))
            return
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        assert noise.ndim == 3 and noise.shape[0] == self.B
        check(self.lib.azb_set_root_noise(self.h, noise.ctypes.data_as(C.c_void_p), noise.shape[1], noise.shape[2]))

    # ---- queues ---------------------------------------------------------------
    def sample_count(self, stream=None):
        n = C.c_int64()
        check(self.lib.azb_sample_count(self.h, C.byref(n), self._stream(stream)))
        return n.value

    def games_played(self, stream=None):
        n = C.c_int64()
        check(self.lib.azb_games_played(self.h, C.byref(n), self._stream(stream)))
        return n.value

    def drain_samples(self, stream=None):
        """-> (obs [n,C,H,W], pi [n,A], z [n,3], slot [n]) as numpy arrays."""
        n = self.sample_count(stream)
        obs = np.empty((n,) + self.obs_shape, dtype=np.float32)
        pi = np.empty((n, self.A), dtype=np.float32)
        z = np.empty((n, 3), dtype=np.float32)
        slot = np.empty(n, dtype=np.int32)
        got = C.c_int64()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.azb_drain_samples(self.h, vp(obs), vp(pi), vp(z), vp(slot), n, C.byref(got), self._stream(stream)))
        assert got.value == n
        return obs, pi, z, slot

    def drain_samples_into(self, obs, pi, z, stream=None):
        """Drain into caller tensors (pinned host or CUDA); returns the count."""
        got = C.c_int64()
        cap = min(obs.shape[0], pi.shape[0], z.shape[0])
        fn = self.lib.azb_drain_samples_device if obs.is_cuda else self.lib.azb_drain_samples
        check(fn(self.h, C.c_void_p(obs.data_ptr()), C.c_void_p(pi.data_ptr()), C.c_void_p(z.data_ptr()), None,
                 cap, C.byref(got), self._stream(stream)))
        return got.value

    def drain_results(self, stream=None):
        """-> (slot [n], turns [n], winstate [n,3] uint8)"""
        cap = 1 << 20
        slot = np.empty(cap, dtype=np.int32)
        turns = np.empty(cap, dtype=np.int32)
        win = np.empty((cap, 3), dtype=np.uint8)
        got = C.c_int64()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.azb_drain_results(self.h, vp(slot), vp(turns), vp(win), cap, C.byref(got), self._stream(stream)))
        n = got.value
        return slot[:n].copy(), turns[:n].copy(), win[:n].copy()

    # ---- introspection ----------------------------------------------------------
    def root_counts(self, stream=None):
        out = np.zeros((self.B, self.A), dtype=np.int32)
        check(self.lib.azb_root_counts(self.h, out.ctypes.data_as(C.c_void_p), self._stream(stream)))
        return out

    def last_actions(self, stream=None):
        out = np.zeros(self.B, dtype=np.int32)
        check(self.lib.azb_game_info(self.h, out.ctypes.data_as(C.c_void_p), None, self._stream(stream)))
        return out

    def turns(self, stream=None):
        out = np.zeros(self.B, dtype=np.int32)
        check(self.lib.azb_game_info(self.h, None, out.ctypes.data_as(C.c_void_p), self._stream(stream)))
        return out

    def boards(self, stream=None):
        out = np.zeros((self.B, self.ncells), dtype=np.int8)
        check(self.lib.azb_boards(self.h, out.ctypes.data_as(C.c_void_p), self._stream(stream)))
        return out

    def tree_dump(self, slot, max_rows=1 << 20, stream=None):
        rows = np.zeros((max_rows, 10), dtype=np.float64)
        got = C.c_int64()
        check(self.lib.azb_tree_dump(self.h, int(slot), rows.ctypes.data_as(C.c_void_p), max_rows, C.byref(got),
                                     self._stream(stream)))
        return rows[:got.value].copy()

    def stats(self, stream=None):
        st = AzbStats()
        check(self.lib.azb_stats_get(self.h, C.byref(st), self._stream(stream)))
        return {n: getattr(st, n) for n, _ in AzbStats._fields_}

    def check_errors(self, stream=None):
        check(self.lib.azb_check_errors(self.h, self._stream(stream)))
