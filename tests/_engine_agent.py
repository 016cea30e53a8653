"""Adapter giving the CUDA engine the same lock-step test surface as
_orc.OracleAgent / _refdriver.RefAgent (generateBatch / processBatch /
playMoves with numpy I/O).  Everything goes through the C ABI."""
import numpy as np
import torch

from azb200 import SelfPlayEngine


class EngineAgent:
    def __init__(self, game="connect4", num_slots=1, rng="mt19937", fused_step_sims=0, **kw):
        self.eng = SelfPlayEngine(game=game, num_games=num_slots, rng=rng, **kw)
        # fused_step_sims = N > 0: processBatch of simulation k < N-1 and generateBatch of k+1 run as ONE launch
        # (azb_expand_backup_select); rounds must then have exactly N simulations
        self._fs, self._k = int(fused_step_sims), 0
        self.B, self.A = self.eng.B, self.eng.A
        self.obs_shape = self.eng.obs_shape

    def set_root_noise(self, noise):
        self.eng.set_root_noise(noise)

    def generateBatch(self):
        if not (self._fs and self._k > 0):
            self.eng.select()
        return self.eng.obs.cpu().numpy()

    def processBatch(self, policy, value):
        self.eng.policy.copy_(torch.from_numpy(np.ascontiguousarray(policy, dtype=np.float32)))
        self.eng.value.copy_(torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)))
        if self._fs and self._k + 1 < self._fs:
            self.eng.expand_backup_select()
            self._k += 1
        else:
            self.eng.expand_backup()
            self._k = 0

    def playMoves(self, fast=False):
        assert self._k == 0
        self.eng.play_moves(fast)
        self.eng.check_errors()

    def root_counts(self):
        return self.eng.root_counts()

    def last_actions(self):
        return self.eng.last_actions()

    def turns(self):
        return self.eng.turns()

    def boards(self):
        return self.eng.boards()

    def stats(self):
        return self.eng.stats()

    def samples(self):
        if not hasattr(self, "_samples"):
            self._samples = [np.zeros((0,) + self.obs_shape, np.float32), np.zeros((0, self.A), np.float32),
                             np.zeros((0, 3), np.float32), np.zeros(0, np.int32)]
        new = self.eng.drain_samples()
        self._samples = [np.concatenate([a, b]) for a, b in zip(self._samples, new)]
        return tuple(self._samples)

    def results(self):
        if not hasattr(self, "_results"):
            self._results = [np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3), np.uint8)]
        new = self.eng.drain_results()
        self._results = [np.concatenate([a, b]) for a, b in zip(self._results, new)]
        return tuple(self._results)

    def tree_dump(self, slot):
        return self.eng.tree_dump(slot)


class ArenaEngineAgent:
    """The engine in arena mode behind the lock-step surface of the oracle's arena agent: game i = slots 2i / 2i+1
    (trees of env player 0 / 1); rows of the idle tree are dropped / left untouched."""

    def __init__(self, game="connect4", num_slots=1, rng="mt19937", mt_seeds=None, arena_temp=0.25,
                 player_to_index=None, **kw):
        if mt_seeds is not None:
            mt_seeds = np.repeat(np.asarray(mt_seeds, dtype=np.uint32), 2)     # the pair shares the even slot's stream
        self.eng = SelfPlayEngine(game=game, num_games=2 * num_slots, rng=rng, mt_seeds=mt_seeds, arena=True,
                                  temps=np.full(1, arena_temp, dtype=np.float64), **kw)
        self.B, self.A = num_slots, self.eng.A
        self.obs_shape = self.eng.obs_shape
        self.player_to_index = list(player_to_index or [0, 1])

    def _active(self):
        pl = self.eng.arena_players().cpu().numpy().reshape(self.B, 2)
        assert ((pl >= 0).sum(1) <= 1).all()
        return pl

    def players(self):
        pl = self._active()
        assert ((pl >= 0).sum(1) == 1).all(), "a finished game has no player to move"
        return pl.max(1)

    def models(self):
        return np.asarray(self.player_to_index, dtype=np.int32)[self.players()]

    def generateBatch(self):
        self.eng.select()
        obs = self.eng.obs.cpu().numpy().reshape((self.B, 2) + self.obs_shape)
        self._rows = self.players()
        return obs[np.arange(self.B), self._rows]

    def processBatch(self, policy, value):
        p = self.eng.policy.cpu().numpy().reshape(self.B, 2, self.A)
        v = self.eng.value.cpu().numpy().reshape(self.B, 2, 3)
        p[np.arange(self.B), self._rows] = policy
        v[np.arange(self.B), self._rows] = value
        self.eng.policy.copy_(torch.from_numpy(p.reshape(2 * self.B, self.A)))
        self.eng.value.copy_(torch.from_numpy(v.reshape(2 * self.B, 3)))
        self.eng.expand_backup()

    def playMoves(self, fast=False):
        self.eng.play_moves(fast)
        self.eng.check_errors()

    def root_counts(self):
        c = self.eng.root_counts().reshape(self.B, 2, self.A)
        return c[np.arange(self.B), self.players()]

    def last_actions(self):
        return self.eng.last_actions().reshape(self.B, 2)[:, 0]

    def turns(self):
        t = self.eng.turns().reshape(self.B, 2)
        assert (t[:, 0] == t[:, 1]).all()
        return t[:, 0]

    def stats(self):
        return self.eng.stats()

    def samples(self):
        return self.eng.drain_samples()

    def results(self):
        if not hasattr(self, "_results"):
            self._results = [np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3), np.uint8)]
        new = self.eng.drain_results()
        self._results = [np.concatenate([a, b]) for a, b in zip(self._results, new)]
        return tuple(self._results)
