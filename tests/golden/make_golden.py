#!/usr/bin/env python
"""Generates the golden traces under tests/golden/ by running the COMPILED
REFERENCE (oracle/_ref, built by oracle/build_ref.py from /root/reference).
Run in the build container only:  python tests/golden/make_golden.py

Each .npz holds, for B game slots driven in lock-step through the reference's
own SelfPlayAgent.generateBatch / processBatch / playMoves:
  counts  [rounds, B, A]  MCTS.counts before every move
  actions [rounds, B]     sampled actions
  turns   [rounds, B]     Game.turns after the move
  s_obs / s_pi / s_z / s_slot   the output_queue in emission order
  r_slot / r_turns / r_win      the result_queue
plus the configuration needed to replay it (seeds, sims, flags, noise table).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _refdriver  # noqa: E402
from _fakenn import FakeNN  # noqa: E402
from _lockstep import run_trace  # noqa: E402

CASES = {
    # unmodified reference: no root temperature, no noise, warmup constants
    "c4_warmup_unmodified": dict(game="connect4", B=3, seeds=[11, 12, 13], rounds=50, sims=25, nn=None,
                                 add_root_temp=False, add_root_noise=False, det_pow=False),
    # fake NN, root temperature with the deterministic pow, fed Dirichlet noise
    "c4_nn_temp_noise": dict(game="connect4", B=4, seeds=[21, 22, 23, 24], rounds=45, sims=30, nn=1234,
                             add_root_temp=True, add_root_noise=True, det_pow=True),
    # fast moves interleaved, no symmetries, tree reset threshold, quota
    "c4_nn_fast_quota": dict(game="connect4", B=4, seeds=[31, 32, 33, 34], rounds=200, sims=10, nn=77,
                             add_root_temp=True, add_root_noise=False, det_pow=True, symmetric=False,
                             reset_threshold=5, quota=6, fast_pattern=[0, 1, 1, 0, 1]),
    # brandubh: unmodified reference (warmup constants), 8-fold symmetric samples, draw at 100 plies
    "tafl_warmup_unmodified": dict(game="brandubh", B=2, seeds=[41, 42], rounds=110, sims=12, nn=None,
                                   add_root_temp=False, add_root_noise=False, det_pow=False),
    "tafl_nn_temp_noise": dict(game="brandubh", B=2, seeds=[51, 52], rounds=60, sims=20, nn=99,
                               add_root_temp=True, add_root_noise=True, det_pow=True),
    # SelfPlayAgent(_is_arena=True): one tree per player, two networks (model of env player p = player_to_index[p]),
    # arenaTemp, tree reset threshold, quota (deterministic pow: probs uses ** (1 / arenaTemp))
    "c4_arena": dict(game="connect4", B=4, seeds=[61, 62, 63, 64], rounds=200, sims=12, nn=[101, 102], arena=True,
                     arena_temp=0.25, player_to_index=[1, 0], add_root_temp=False, det_pow=True, reset_threshold=6, quota=9),
    # non-default search hyper-parameters (cpuct, fpu_reduction, noise fraction, root temperature, start temperature 2
    # on the halving schedule): nothing on the path may have the DEFAULT_ARGS values baked in
    "c4_hyper": dict(game="connect4", B=4, seeds=[81, 82, 83, 84], rounds=45, sims=40, nn=4321,
                     add_root_temp=True, add_root_noise=True, det_pow=True,
                     hyper=dict(cpuct=2.5, fpu_reduction=0.45, root_noise_frac=0.3, root_policy_temp=1.6, start_temp=2.0)),
    "tafl_hyper_fast": dict(game="brandubh", B=3, seeds=[91, 92, 93], rounds=140, sims=10, nn=555,
                            add_root_temp=True, add_root_noise=False, det_pow=True, symmetric=False, reset_threshold=7,
                            quota=4, fast_pattern=[1, 0, 0, 1],
                            hyper=dict(cpuct=0.8, fpu_reduction=0.0, root_noise_frac=0.1, root_policy_temp=0.9, start_temp=1.0)),
    # hnefatafl 11x11: fixtures for the oracle now and for the engine once it serves 121-cell boards --
    # warmup constants (unmodified reference) and a network with root temperature + fed noise, fast moves interleaved
    "hnefatafl_warmup_unmodified": dict(game="hnefatafl", B=2, seeds=[101, 102], rounds=40, sims=8, nn=None,
                                        add_root_temp=False, add_root_noise=False, det_pow=False),
    "hnefatafl_nn_temp_noise": dict(game="hnefatafl", B=2, seeds=[111, 112], rounds=30, sims=10, nn=777,
                                    add_root_temp=True, add_root_noise=True, det_pow=True, fast_pattern=[0, 0, 1]),
    "tafl_arena": dict(game="brandubh", B=2, seeds=[71, 72], rounds=45, sims=8, nn=[111, 112], arena=True,
                       arena_temp=0.5, player_to_index=[0, 1], add_root_temp=False, det_pow=True),
}
GAME_DIMS = {"connect4": (4 * 6 * 7, 7), "brandubh": (5 * 7 * 7, 588), "hnefatafl": (5 * 11 * 11, 2420)}


def make(name, c):
    obs_size, A = GAME_DIMS[c["game"]]
    noise = None
    if c.get("add_root_noise"):
        rs = np.random.RandomState(1000 + len(name))
        noise = rs.dirichlet([10.83 / 7] * 7, size=(c["B"], 24)).astype(np.float32)
        if A > 7:
            noise = rs.dirichlet([10.83 / 40] * 96, size=(c["B"], 8)).astype(np.float32)
        if A > 588:                       # hnefatafl: up to ~150 legal moves at a root
            with np.errstate(under="ignore"):       # the reference runs under np.seterr(all='raise') (MCTS.pyx:23)
                noise = rs.dirichlet([10.83 / 116] * 256, size=(c["B"], 8)).astype(np.float32)
    arena = bool(c.get("arena"))
    hyper = dict(cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1, start_temp=1.0)
    hyper.update(c.get("hyper", {}))
    ref = _refdriver.RefAgent(c["game"], c["B"], mt_seeds=c["seeds"], add_root_temp=c["add_root_temp"], **hyper,
                              add_root_noise=c.get("add_root_noise", False), det_pow=c["det_pow"], noise=noise,
                              symmetric_samples=c.get("symmetric", True),
                              mcts_reset_threshold=c.get("reset_threshold"),
                              games_per_iteration=c.get("quota", 1 << 40), arena=arena,
                              arena_temp=c.get("arena_temp", 0.25), player_to_index=c.get("player_to_index"))
    if arena:
        from _fakenn import ArenaNN
        nn = ArenaNN(ref, [FakeNN(obs_size, A, seed=sd, sharp=3.0 if A == 7 else 1.0) for sd in c["nn"]])
    else:
        nn = FakeNN(obs_size, A, seed=c["nn"], sharp=3.0 if A == 7 else 1.0) if c["nn"] is not None else None
    tr = run_trace(ref, nn, c["rounds"], c["sims"], fast_pattern=c.get("fast_pattern"), until_games=c.get("quota"))
    s_obs, s_pi, s_z, s_slot = ref.samples()
    r_slot, r_turns, r_win = ref.results()
    out = dict(counts=np.stack([t["counts"] for t in tr]), actions=np.stack([t["actions"] for t in tr]),
               turns=np.stack([t["turns"] for t in tr]), s_obs=s_obs.astype(np.float16 if False else np.float32),
               s_pi=s_pi, s_z=s_z, s_slot=s_slot, r_slot=r_slot, r_turns=r_turns, r_win=r_win,
               seeds=np.asarray(c["seeds"]), sims=c["sims"],
               nn_seed=-1 if (c["nn"] is None or arena) else c["nn"], arena=arena, arena_temp=c.get("arena_temp", 0.25),
               arena_nn_seeds=np.asarray(c["nn"] if arena else [0, 0]), player_to_index=np.asarray(c.get("player_to_index") or [0, 1]),
               add_root_temp=c["add_root_temp"], add_root_noise=c.get("add_root_noise", False),
               symmetric=c.get("symmetric", True), reset_threshold=c.get("reset_threshold") or 0,
               quota=c.get("quota", 0), fast_pattern=np.asarray(c.get("fast_pattern", [0])),
               noise=noise if noise is not None else np.zeros((0, 0, 0), np.float32), game=c["game"],
               **{"hyper_" + k: np.float64(v) for k, v in hyper.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rounds", len(tr), "samples", len(s_obs), "results", len(r_slot))


if __name__ == "__main__":
    assert _refdriver.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    only = sys.argv[1:]
    for name, c in CASES.items():
        if not only or name in only:
            make(name, c)
