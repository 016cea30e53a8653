"""Adapter giving the CUDA engine the same lock-step test surface as
_orc.OracleAgent / _refdriver.RefAgent (generateBatch / processBatch /
playMoves with numpy I/O).  Everything goes through the C ABI."""
import numpy as np
import torch

from azb200 import SelfPlayEngine


class EngineAgent:
    def __init__(self, game="connect4", num_slots=1, rng="mt19937", **kw):
        self.eng = SelfPlayEngine(game=game, num_games=num_slots, rng=rng, **kw)
        self.B, self.A = self.eng.B, self.eng.A
        self.obs_shape = self.eng.obs_shape

    def set_root_noise(self, noise):
        self.eng.set_root_noise(noise)

    def generateBatch(self):
        self.eng.select()
        return self.eng.obs.cpu().numpy()

    def processBatch(self, policy, value):
        self.eng.policy.copy_(torch.from_numpy(np.ascontiguousarray(policy, dtype=np.float32)))
        self.eng.value.copy_(torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)))
        self.eng.expand_backup()

    def playMoves(self, fast=False):
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
