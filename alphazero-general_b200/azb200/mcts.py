"""Single-tree search API on the engine: the public surface of alphazero/MCTS.pyx's ``MCTS`` class
(:120-195, :297-345) that GenericPlayers.MCTSPlayer / RawMCTSPlayer (GenericPlayers.py:100-200), the evaluator and the
tree plot use -- ``search / raw_search / update_root / counts / probs / best_action / value / reset`` -- served by a
one-slot SelfPlayEngine:

    mcts = MCTS(args)                                   # args: cpuct, fpu_reduction, root_noise_frac, root_policy_temp
    mcts.search(state, nn, args.numMCTSSims, add_root_noise, add_root_temp)
    policy = mcts.probs(state, temp)
    mcts.update_root(state, action)                     # keeps the chosen subtree, as the reference does

``state`` is the reference's Game object (Connect4 or brandubh plugin); it is only read (board cells, turns,
observation()).  The engine slot is put on the caller's position whenever it is not already there (azb_set_state) and
follows ``update_root`` with azb_force_move, so the subtree is reused exactly like MCTS._root.  With ``rng="mt19937"``
and ``np.random.seed(s)`` <-> ``seed=s`` the child orders -- and therefore visit counts -- equal the reference's
(tests/test_mcts_api.py).  This is the API-parity surface, not a throughput path: one tree per engine, one leaf per
launch; self-play and the arena use the batched drivers.
"""
import numpy as np

from .engine import SelfPlayEngine
from .selfplay import game_name


def _cells(gs):
    """Reference cell codes of a Game object: Connect4 Board.pieces, tafl Board._state (both row-major)."""
    b = gs._board
    a = getattr(b, "pieces", None)
    if a is None:
        a = getattr(b, "_state")
    return np.asarray(a, dtype=np.int8).ravel()


class MCTS:
    def __init__(self, args, device=0, rng="mt19937", seed=0):
        g = lambda k, d: (args[k] if k in args else d)
        self.args = args
        self.root_noise_frac = float(g("root_noise_frac", 0.1))
        self.root_temp = float(g("root_policy_temp", 1.1))
        self.fpu_reduction = float(g("fpu_reduction", 0.2))
        self.cpuct = float(g("cpuct", 1.25))
        self.max_depth = 0
        self._device, self._rng, self._seed = device, rng, int(seed)
        self._eng = None
        self._pos = None              # (cells bytes, turns) the engine slot is on
        self._sims_cap = int(g("numMCTSSims", 100))

    # ---- engine plumbing ---------------------------------------------------------------------------------
    def _engine(self, gs):
        if self._eng is None:
            kw = dict(game=game_name(type(gs)), num_games=1, device=self._device, rng=self._rng, cpuct=self.cpuct,
                      fpu_reduction=self.fpu_reduction, root_noise_frac=self.root_noise_frac,
                      root_policy_temp=self.root_temp, max_sims_per_move=max(self._sims_cap, 1), games_per_iteration=0)
            if self._rng == "mt19937":
                kw["mt_seeds"] = [self._seed]
            else:
                kw["seed"] = self._seed
            self._eng = SelfPlayEngine(**kw)
        return self._eng

    def _sync(self, gs):
        eng = self._engine(gs)
        pos = (_cells(gs).tobytes(), int(gs.turns))
        if pos != self._pos:                                     # a new position: empty tree (MCTS.reset semantics)
            eng.set_state(0, np.frombuffer(pos[0], dtype=np.int8), pos[1])
            self._pos = pos
        return eng

    def reset(self):
        """MCTS.reset (MCTS.pyx:144-150)."""
        self._pos = None
        self.max_depth = 0

    # ---- search ---------------------------------------------------------------------------------------------
    def _run(self, gs, sims, add_root_noise, add_root_temp, evaluate):
        import torch
        eng = self._sync(gs)
        eng.set_root_flags(add_root_noise, add_root_temp)
        self.max_depth = 0
        d0 = eng.stats()["sum_depth"]
        for _ in range(int(sims)):
            eng.select()
            p, v = evaluate(eng.obs[0].cpu().numpy())
            eng.policy[0].copy_(torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32)))
            eng.value[0].copy_(torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)))
            eng.expand_backup()
            d1 = eng.stats()["sum_depth"]
            self.max_depth = max(self.max_depth, int(d1 - d0))
            d0 = d1
        eng.check_errors()

    def search(self, gs, nn, sims, add_root_noise, add_root_temp):
        """MCTS.search (MCTS.pyx:154-162): ``nn(observation) -> (policy, value)``."""
        self._run(gs, sims, add_root_noise, add_root_temp, lambda obs: nn(obs))

    def raw_search(self, gs, sims, add_root_noise, add_root_temp):
        """MCTS.raw_search (MCTS.pyx:164-173): constant policy of ones, zero value."""
        p = np.full(gs.action_size(), 1, dtype=np.float32)
        v = np.zeros(gs.num_players() + 1, dtype=np.float32)
        self._run(gs, sims, add_root_noise, add_root_temp, lambda obs: (p, v))

    def update_root(self, gs, a):
        """MCTS.update_root (MCTS.pyx:185-195); `gs` is the state before the move."""
        eng = self._sync(gs)
        valid = np.asarray(gs.valid_moves()).astype(bool)
        if not (0 <= int(a) < len(valid)) or not valid[int(a)]:
            raise ValueError(f"Invalid action while updating root: {a}")
        eng.force_move(0, int(a))
        eng.check_errors()
        self._pos = None
        nxt = gs.clone()
        nxt.play_action(int(a))
        self._pos = (_cells(nxt).tobytes(), int(nxt.turns))

    # ---- read-outs ------------------------------------------------------------------------------------------
    def counts(self, gs):
        """MCTS.counts (MCTS.pyx:297-303)."""
        return self._sync(gs).root_counts()[0].astype(np.int32)

    def best_action(self, gs):
        return int(np.argmax(self.counts(gs)))

    def probs(self, gs, temp=1.0):
        """MCTS.probs (MCTS.pyx:308-327), NumPy float32 arithmetic as the reference's."""
        counts = np.array(self.counts(gs), dtype=np.float32)
        if temp == 0:
            probs = np.zeros_like(counts)
            probs[int(np.argmax(counts))] = 1
            return probs
        try:
            with np.errstate(all="raise"):
                probs = (counts / np.sum(counts)) ** np.float32(1.0 / float(np.float32(temp)))
                probs /= np.sum(probs)
            return probs
        except (OverflowError, FloatingPointError):
            probs = np.zeros_like(counts)
            probs[int(np.argmax(counts))] = 1
            return probs

    def _root_children(self):
        rows = self._eng.tree_dump(0)
        return rows[rows[:, 0] == 1] if len(rows) else rows       # depth-1 rows: (depth, a, n, q, v, p, player, e0..e2)

    def value(self, average=False):
        """MCTS.value (MCTS.pyx:329-345): max (or mean over all children) of the visited root children's q."""
        ch = self._root_children()
        if not len(ch):
            return 0.0
        vis = ch[ch[:, 2] > 0]
        if average:
            return float(np.float32(sum(float(q) for q in vis[:, 3]) / len(ch)))
        value = np.float32(0)
        for q in vis[:, 3]:
            if np.float32(q) > value:
                value = np.float32(q)
        return float(value)
