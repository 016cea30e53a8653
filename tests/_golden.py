"""Replay helper for the golden traces in tests/golden/ (made from the compiled
reference by tests/golden/make_golden.py)."""
import os

import numpy as np

import _orc
from _fakenn import ArenaNN, FakeNN
from _lockstep import run_trace

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAME_IDS = {"connect4": _orc.GAME_CONNECT4, "brandubh": _orc.GAME_BRANDUBH, "hnefatafl": _orc.GAME_HNEFATAFL}
DIMS = {"connect4": (4 * 6 * 7, 7, 42), "brandubh": (5 * 7 * 7, 588, None), "hnefatafl": (5 * 11 * 11, 2420, None)}
ENGINE_GAMES = ("connect4", "brandubh", "hnefatafl")      # what libazb200.so serves


def cases(arena=False, engine=False):
    names = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and (f[:-4].endswith("_arena") == arena))
    if engine:
        names = [n for n in names if any(n.startswith(g if g != "brandubh" else "tafl") or n.startswith(g[:2]) for g in ENGINE_GAMES)]
    return names


def is_arena(g):
    return "arena" in g.files and bool(g["arena"])


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def agent_kwargs(g):
    game = str(g["game"])
    _, _, max_turns = DIMS[game]
    if is_arena(g):
        kw = dict(mcts_reset_threshold=int(g["reset_threshold"]), games_per_iteration=int(g["quota"]) or (1 << 40),
                  arena_temp=float(g["arena_temp"]), player_to_index=g["player_to_index"].tolist())
        return game, kw
    hp = lambda k, d: float(g["hyper_" + k]) if ("hyper_" + k) in g.files else d       # older files: DEFAULT_ARGS
    kw = dict(add_root_temp=bool(g["add_root_temp"]), add_root_noise=bool(g["add_root_noise"]),
              symmetric_samples=bool(g["symmetric"]), mcts_reset_threshold=int(g["reset_threshold"]),
              games_per_iteration=int(g["quota"]) or (1 << 40),
              cpuct=hp("cpuct", 1.25), fpu_reduction=hp("fpu_reduction", 0.2), root_noise_frac=hp("root_noise_frac", 0.1),
              root_policy_temp=hp("root_policy_temp", 1.1),
              temps=_orc.temp_table(_orc.default_temp_scaling, hp("start_temp", 1.0), max_turns))
    return game, kw


def replay(agent, g):
    game = str(g["game"])
    obs_size, A, _ = DIMS[game]
    if bool(g["add_root_noise"]):
        agent.set_root_noise(g["noise"])
    if is_arena(g):
        nn = ArenaNN(agent, [FakeNN(obs_size, A, seed=int(sd), sharp=3.0 if A == 7 else 1.0) for sd in g["arena_nn_seeds"]])
    else:
        nn = FakeNN(obs_size, A, seed=int(g["nn_seed"]), sharp=3.0 if A == 7 else 1.0) if int(g["nn_seed"]) >= 0 else None
    pat = g["fast_pattern"].tolist()
    return run_trace(agent, nn, len(g["counts"]), int(g["sims"]), fast_pattern=pat if any(pat) else None,
                     until_games=int(g["quota"]) or None)


def check(agent, g, trace):
    assert len(trace) == len(g["counts"])
    for r, t in enumerate(trace):
        assert np.array_equal(t["counts"], g["counts"][r]), f"visit counts differ at round {r}"
        assert np.array_equal(t["actions"], g["actions"][r]), f"actions differ at round {r}"
        assert np.array_equal(t["turns"], g["turns"][r]), f"turns differ at round {r}"
    s = agent.samples()
    for name, got in zip(("s_obs", "s_pi", "s_z", "s_slot"), s):
        assert got.shape == g[name].shape, (name, got.shape, g[name].shape)
        assert np.array_equal(got, g[name]), f"{name} differ"
    for name, got in zip(("r_slot", "r_turns", "r_win"), agent.results()):
        assert np.array_equal(got, g[name]), f"{name} differ"
