"""Deterministic stand-in for NNetWrapper.process used by the parity tests:
a fixed random two-layer net evaluated in float64 on the host and rounded to
float32, so the engine under test and the oracle are fed bit-identical
(policy, value) rows for bit-identical observations.  Test infrastructure."""
import numpy as np


class FakeNN:
    def __init__(self, obs_size, action_size, seed=1234, sharp=3.0):
        rng = np.random.RandomState(seed)
        self.w1 = rng.standard_normal((obs_size, 32)) / np.sqrt(obs_size) * 4.0
        self.wp = rng.standard_normal((32, action_size)) * sharp
        self.wv = rng.standard_normal((32, 3)) * 1.5

    def __call__(self, obs):
        x = np.asarray(obs, dtype=np.float64).reshape(len(obs), -1)
        h = np.tanh(np.stack([r @ self.w1 for r in x]))      # row-wise: independent of batch size
        lp = np.stack([r @ self.wp for r in h])
        lv = np.stack([r @ self.wv for r in h])
        # importing the compiled reference switches NumPy to np.seterr(all='raise') (MCTS.pyx:23); a probability that is
        # subnormal or zero in float32 is a perfectly good network answer
        with np.errstate(under="ignore"):
            p = np.exp(lp - lp.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
            v = np.exp(lv - lv.max(1, keepdims=True)); v /= v.sum(1, keepdims=True)
            return p.astype(np.float32), v.astype(np.float32)


def warmup_outputs(batch, action_size):
    """SelfPlayAgent warmup constants (SelfPlayAgent.pyx:48-52): float32(1/A), float32(1/3)."""
    p = np.full((batch, action_size), np.float32(1.0 / action_size), dtype=np.float32)
    v = np.full((batch, 3), np.float32(1.0 / 3), dtype=np.float32)
    return p, v


class ArenaNN:
    """Arena.play_games' server loop (Arena.pyx:262-275) for the lock-step drivers: every row is evaluated by the
    model of the player to move in that game, `agent.models()` says which."""

    def __init__(self, agent, nets):
        self.agent, self.nets = agent, nets

    def __call__(self, obs):
        models = self.agent.models()
        outs = [net(obs) for net in self.nets]
        p = np.stack([outs[m][0][i] for i, m in enumerate(models)])
        v = np.stack([outs[m][1][i] for i, m in enumerate(models)])
        return p, v
