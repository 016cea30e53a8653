"""In-process driver of the compiled reference (oracle/_ref).  Test infrastructure.

The reference's own ``SelfPlayAgent`` (alphazero/SelfPlayAgent.pyx) is built
with batch size 1 per game slot and driven through its own
``generateBatch / processBatch / playMoves`` with list-backed queues and no-op
events -- ``run()`` is not used because it reseeds from OS entropy
(SelfPlayAgent.pyx:81).  Each slot owns a legacy MT19937 state that is swapped
into ``np.random`` around every call, so slot i behaves exactly like a
reference agent seeded with ``np.random.seed(seed_i)``.

Two optional patches of the reference, each limited to one operation and both
documented in DESIGN.md:
  * ``det_pow``: float32 ``**`` (MCTS.pyx:250,320) is evaluated as the correctly
    rounded power through libm's double pow, because NumPy's float32 power is
    SIMD-dispatch dependent (SVML vs libm differ by 1 ulp on ~20% of inputs).
  * ``noise``: ``np.random.dirichlet`` (MCTS.pyx:199) returns host-fed vectors
    instead of consuming the MT stream (numpy's gamma sampler cannot be
    reproduced bit-for-bit on a GPU).
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref
        return build_ref.built()
    finally:
        sys.path.pop(0)


def _import_ref():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torch  # noqa: F401  (SelfPlayAgent imports torch.multiprocessing)
    import alphazero.MCTS as ref_mcts
    import alphazero.SelfPlayAgent as ref_spa
    from alphazero.utils import dotdict, default_temp_scaling
    return ref_mcts, ref_spa, dotdict, default_temp_scaling


class _DetPowArray(np.ndarray):
    def __pow__(self, e):
        e32 = float(np.float32(e))
        if e32 == 1.0:
            return np.array(self, dtype=np.float32)
        flat = np.asarray(self, dtype=np.float32).ravel()
        out = np.array([math.pow(float(x), e32) for x in flat], dtype=np.float64)
        return out.astype(np.float32).reshape(self.shape)


class _DetSum(np.float32):
    def __rtruediv__(self, other):
        return (np.asarray(other, dtype=np.float32) / np.float32(self)).view(_DetPowArray)


class _RandomShim:
    def __init__(self, owner):
        self._owner = owner

    def __getattr__(self, name):
        return getattr(np.random, name)

    def dirichlet(self, alpha, size=None):
        return self._owner._next_noise(len(alpha))


class _NpShim:
    """Stands in for the ``np`` global of alphazero.MCTS."""

    def __init__(self, det_pow, feed_noise):
        self._det_pow = det_pow
        self._noise_src = None
        self.random = _RandomShim(self) if feed_noise else np.random

    def __getattr__(self, name):
        return getattr(np, name)

    def asarray(self, x, *a, **k):
        r = np.asarray(x, *a, **k)
        if self._det_pow and r.dtype == np.float32:
            return r.view(_DetPowArray)
        return r

    def sum(self, x, *a, **k):
        r = np.sum(x, *a, **k)
        if self._det_pow and isinstance(r, np.float32):
            return _DetSum(r)
        return r

    def _next_noise(self, n):
        return self._noise_src(n)


class _FakeQueue:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)

    def qsize(self):
        return len(self.items)

    def close(self):
        pass

    def join_thread(self):
        pass


class _FakeEvent:
    def is_set(self):
        return False

    def wait(self):
        pass

    def clear(self):
        pass

    def set(self):
        pass


class _FakeLock:
    def acquire(self):
        pass

    def release(self):
        pass


class _FakeValue:
    """Stands in for mp.Value('i'); shared by all slots (games_played quota)."""

    def __init__(self, v=0):
        self.value = v
        self._lock = _FakeLock()

    def get_lock(self):
        return self._lock


def game_class(name):
    _import_ref()
    if name == "connect4":
        from alphazero.envs.connect4.connect4 import Game
        return Game
    if name in ("brandubh", "hnefatafl"):
        if name == "brandubh":
            from alphazero.envs.brandubh.fastafl import Game as _G
        else:
            from alphazero.envs.hnefatafl.fastafl import Game as _G

        class Game(_G):
            # the reference env lacks these two GameState statics (Game.py:55-63)
            @staticmethod
            def max_turns():
                return None

            @staticmethod
            def has_draw():
                return True

        return Game
    raise KeyError(name)


class RefAgent:
    """B independent reference SelfPlayAgent(batch size 1) instances in lock-step."""

    def __init__(self, game="connect4", num_slots=1, mt_seeds=None, cpuct=1.25, fpu_reduction=0.2,
                 root_noise_frac=0.1, root_policy_temp=1.1, add_root_noise=False, add_root_temp=False,
                 symmetric_samples=True, mcts_reset_threshold=None, games_per_iteration=1 << 40,
                 temp_scaling_fn=None, start_temp=1, det_pow=False, noise=None, arena=False, arena_temp=0.25,
                 player_to_index=None):
        import torch
        self.ref_mcts, self.ref_spa, dotdict, default_temp_scaling = _import_ref()
        self.game_cls = game_class(game)
        self.B = num_slots
        self.A = self.game_cls.action_size()
        self.obs_shape = tuple(self.game_cls.observation_size())
        self.args = dotdict(dict(
            startTemp=start_temp, temp_scaling_fn=temp_scaling_fn or default_temp_scaling,
            root_noise_frac=root_noise_frac, root_policy_temp=root_policy_temp, min_discount=1,
            fpu_reduction=fpu_reduction, cpuct=cpuct, _num_players=self.game_cls.num_players(),
            add_root_noise=add_root_noise, add_root_temp=add_root_temp,
            symmetricSamples=symmetric_samples, mctsResetThreshold=mcts_reset_threshold,
            gamesPerIteration=games_per_iteration, numMCTSSims=1, numFastSims=1, probFastSim=0.0,
            arenaTemp=arena_temp,
        ))
        self.arena = arena
        self.shim = _NpShim(det_pow, noise is not None) if (det_pow or noise is not None) else None
        self.noise = None if noise is None else np.asarray(noise, dtype=np.float32)
        self.noise_event = [0] * num_slots
        self._cur_slot = 0
        if self.shim is not None:
            self.shim._noise_src = self._noise_for
        self.games_played = _FakeValue(0)
        self.complete = _FakeValue(0)
        self.out_q = [_FakeQueue() for _ in range(num_slots)]
        self.res_q = [_FakeQueue() for _ in range(num_slots)]
        self.sample_order = []           # (slot, index in that slot's queue) in emission order
        self.result_order = []
        self.agents, self.states = [], []
        self.bt, self.pt, self.vt = [], [], []
        mt_seeds = list(range(num_slots)) if mt_seeds is None else list(mt_seeds)
        for i in range(num_slots):
            bt = torch.zeros((1,) + self.obs_shape)
            pt = torch.zeros((1, self.A))
            vt = torch.zeros((1, 3))
            np.random.seed(int(mt_seeds[i]))
            if arena:
                # Arena.play_games (Arena.pyx:236-258): batch_tensor is a per-player list, the batches travel through
                # output_queue; the constructor shuffles player_to_index (SelfPlayAgent.pyx:44-47) -- fixed here, and
                # the slot's stream is seeded after construction so that the shuffle does not consume from it
                ag = self.ref_spa.SelfPlayAgent(i, self.game_cls, _FakeQueue(), _FakeEvent(), [[], []], pt, vt,
                                                self.out_q[i], self.res_q[i], self.complete, self.games_played,
                                                _FakeEvent(), _FakeEvent(), self.args, _is_arena=True)
                ag.player_to_index = list(player_to_index or [0, 1])
                np.random.seed(int(mt_seeds[i]))
            else:
                ag = self.ref_spa.SelfPlayAgent(i, self.game_cls, _FakeQueue(), _FakeEvent(), bt, pt, vt,
                                                self.out_q[i], self.res_q[i], self.complete, self.games_played,
                                                _FakeEvent(), _FakeEvent(), self.args)
            self.states.append(np.random.get_state())
            self.agents.append(ag)
            self.bt.append(bt); self.pt.append(pt); self.vt.append(vt)
        self.last_action = [-1] * num_slots

    def _noise_for(self, n):
        s = self._cur_slot
        e = self.noise_event[s]
        self.noise_event[s] += 1
        return np.asarray(self.noise[s, e, :n], dtype=np.float64)

    def _enter(self, i):
        self._cur_slot = i
        np.random.set_state(self.states[i])
        if self.shim is not None:
            self._saved_np = self.ref_mcts.np
            self.ref_mcts.np = self.shim

    def _leave(self, i):
        self.states[i] = np.random.get_state()
        if self.shim is not None:
            self.ref_mcts.np = self._saved_np

    def generateBatch(self):
        obs = np.zeros((self.B,) + self.obs_shape, dtype=np.float32)
        for i, ag in enumerate(self.agents):
            self._enter(i)
            try:
                ag.generateBatch()
            finally:
                self._leave(i)
            if self.arena:
                batch = self.out_q[i].items.pop()                 # [per model: tensor or []]
                model = ag.player_to_index[ag.games[0].player]
                assert [isinstance(b, list) for b in batch] == [m != model for m in range(len(batch))]
                obs[i] = batch[model][0].numpy()
            else:
                obs[i] = self.bt[i][0].numpy()
        return obs

    def models(self):
        """arena: index of the model that evaluates each slot's current leaf (player_to_index[game.player])."""
        return np.asarray([ag.player_to_index[ag.games[0].player] for ag in self.agents], dtype=np.int32)

    def players(self):
        return np.asarray([ag.games[0].player for ag in self.agents], dtype=np.int32)

    def processBatch(self, policy, value):
        import torch
        for i, ag in enumerate(self.agents):
            self.pt[i][0].copy_(torch.from_numpy(np.ascontiguousarray(policy[i], dtype=np.float32)))
            self.vt[i][0].copy_(torch.from_numpy(np.ascontiguousarray(value[i], dtype=np.float32)))
            self._enter(i)
            try:
                ag.processBatch()
            finally:
                self._leave(i)

    def playMoves(self, fast=False):
        for i, ag in enumerate(self.agents):
            ag.fast = fast
            before_s, before_r = (0 if self.arena else self.out_q[i].qsize()), self.res_q[i].qsize()
            turns_before = ag.games[0].turns
            self._enter(i)
            try:
                ag.playMoves()
            finally:
                self._leave(i)
            for k in range(before_s, 0 if self.arena else self.out_q[i].qsize()):
                self.sample_order.append((i, k))
            for k in range(before_r, self.res_q[i].qsize()):
                self.result_order.append((i, k))
            g = ag.games[0]
            if g.turns == turns_before + 1:
                self.last_action[i] = int(g.last_action)
            else:   # game finished and was replaced: the action is on the result's final state
                self.last_action[i] = int(self.res_q[i].items[-1][0].last_action)

    def root_counts(self):
        out = np.zeros((self.B, self.A), dtype=np.int32)
        for i, ag in enumerate(self.agents):
            out[i] = np.asarray(ag._mcts(0).counts(ag.games[0]))
        return out

    def last_actions(self):
        return np.asarray(self.last_action, dtype=np.int32)

    def turns(self):
        return np.asarray([ag.games[0].turns for ag in self.agents], dtype=np.int32)

    def samples(self):
        n = len(self.sample_order)
        obs = np.zeros((n,) + self.obs_shape, dtype=np.float32)
        pi = np.zeros((n, self.A), dtype=np.float32)
        z = np.zeros((n, 3), dtype=np.float32)
        slot = np.zeros(n, dtype=np.int32)
        for j, (i, k) in enumerate(self.sample_order):
            o, p, w = self.out_q[i].items[k]
            obs[j], pi[j], z[j], slot[j] = o, p, w, i
        return obs, pi, z, slot

    def results(self):
        n = len(self.result_order)
        slot = np.zeros(n, dtype=np.int32)
        turns = np.zeros(n, dtype=np.int32)
        win = np.zeros((n, 3), dtype=np.uint8)
        for j, (i, k) in enumerate(self.result_order):
            st, w, _ = self.res_q[i].items[k]
            slot[j], turns[j], win[j] = i, st.turns, w
        return slot, turns, win

    def tree_dump(self, slot, max_nodes=100000):
        """Pre-order dump of a slot's tree: rows (depth, a, n, q, v, p, player, e0,e1,e2)."""
        rows = []

        def rec(nd, d):
            e = [int(x) for x in nd.e] + [0, 0, 0]   # unvisited nodes carry a 2-entry e (MCTS.pyx:62)
            rows.append((d, nd.a, nd.n, nd.q, nd.v, nd.p, nd.player, *e[:3]))
            for c in nd._children:
                if len(rows) < max_nodes:
                    rec(c, d + 1)

        rec(self.agents[slot].mcts[0]._root, 0)
        return np.asarray(rows, dtype=np.float64)

    def stats(self):
        return {"games_played": int(self.games_played.value), "samples": len(self.sample_order),
                "results": len(self.result_order)}
