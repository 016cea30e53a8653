"""Shared lock-step driver for the parity tests: runs any agent exposing
generateBatch / processBatch / playMoves and records a comparable trace."""
import numpy as np

from _fakenn import warmup_outputs


def run_trace(agent, nn, rounds, sims, fast_pattern=None, keep_obs=False, until_games=None):
    """-> list of per-round dicts: counts before the move, actions, turns after."""
    trace = []
    for r in range(rounds):
        obs_sig = []
        for _ in range(sims):
            obs = agent.generateBatch()
            if keep_obs:
                obs_sig.append(obs.copy())
            p, v = nn(obs) if nn is not None else warmup_outputs(len(obs), agent.A)
            agent.processBatch(p, v)
        counts = agent.root_counts().copy()
        fast = bool(fast_pattern[r % len(fast_pattern)]) if fast_pattern else False
        agent.playMoves(fast)
        trace.append(dict(counts=counts, actions=agent.last_actions().copy(), turns=agent.turns().copy(),
                          obs=obs_sig))
        if until_games is not None and agent.stats()["games_played"] >= until_games:
            break
    return trace


def assert_traces_equal(ta, tb, what=""):
    assert len(ta) == len(tb)
    for r, (a, b) in enumerate(zip(ta, tb)):
        for key in ("counts", "actions", "turns"):
            if not np.array_equal(a[key], b[key]):
                bad = np.argwhere(np.asarray(a[key]) != np.asarray(b[key]))[0]
                raise AssertionError(f"{what} round {r}: {key} differ first at {bad}: "
                                     f"{np.asarray(a[key])[tuple(bad)]} vs {np.asarray(b[key])[tuple(bad)]}")
        for oa, ob in zip(a["obs"], b["obs"]):
            assert np.array_equal(oa, ob), f"{what} round {r}: leaf observations differ"


def assert_queues_equal(a, b, what=""):
    for name, x, y in zip(("obs", "pi", "z", "slot"), a.samples(), b.samples()):
        assert x.shape == y.shape, f"{what} samples.{name} shape {x.shape} vs {y.shape}"
        assert np.array_equal(x, y), f"{what} samples.{name} differ"
    for name, x, y in zip(("slot", "turns", "win"), a.results(), b.results()):
        assert np.array_equal(x, y), f"{what} results.{name} differ"
