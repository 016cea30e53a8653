"""Batched arena on the engine (azb200/arena.py) against the oracle's arena mode: the device-resident
play_games and the reference-shaped ArenaAgent driven by the loop body of Arena.play_games (Arena.pyx:262-275)."""
import queue

import numpy as np
import pytest
import torch

import _orc
from _fakenn import ArenaNN, FakeNN

pytestmark = pytest.mark.gpu


class _Args(dict):
    __getattr__ = dict.__getitem__


class _C4Game:
    __module__ = "alphazero.envs.connect4.connect4"

    @staticmethod
    def max_turns():
        return 42

    @staticmethod
    def num_players():
        return 2


class _Model:
    """NNetWrapper.process surface over a FakeNN (evaluated on the host in float64 like the oracle's)."""

    def __init__(self, net):
        self.net = net

    def process(self, batch):
        p, v = self.net(batch.detach().cpu().numpy())
        return torch.from_numpy(p).to(batch.device), torch.from_numpy(v).to(batch.device)


def _oracle_run(B, sims, quota, seed, p2i, nets, base=0):
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, rng_mode=_orc.RNG_PHILOX, seed=seed, game_id_base=base, arena=True,
                           arena_temp=0.25, player_to_index=p2i, games_per_iteration=quota)
    nn = ArenaNN(orc, nets)
    while orc.stats()["games_played"] < quota:
        for _ in range(sims):
            obs = orc.generateBatch()
            orc.processBatch(*nn(obs))
        orc.playMoves(False)
    return orc


def _args(sims, quota):
    return _Args(numMCTSSims=sims, numFastSims=sims, numWarmupSims=sims, gamesPerIteration=quota, arenaTemp=0.25,
                 cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1, add_root_noise=True,
                 add_root_temp=True, symmetricSamples=True, mctsResetThreshold=None, startTemp=1, probFastSim=0.0)


@pytest.mark.parametrize("p2i", [(0, 1), (1, 0)])
def test_play_games_matches_oracle(p2i):
    from azb200.arena import arena_engine, play_games
    B, sims, quota, seed = 16, 8, 40, 3
    nets = [FakeNN(4 * 6 * 7, 7, seed=80), FakeNN(4 * 6 * 7, 7, seed=81)]
    eng = arena_engine(_C4Game, _args(sims, quota), B, rng="philox", seed=seed)
    wins, draws, mean_turns, nsims = play_games(eng, [_Model(n) for n in nets], p2i, sims=sims)
    orc = _oracle_run(B, sims, quota, seed, list(p2i), nets)
    _, turns, win = orc.results()
    turns, win = turns[:quota], win[:quota]
    want = [0, 0]
    for p in range(2):
        want[p2i[p]] += int(win[:, p].sum())
    assert wins == want and draws == int(win[:, 2].sum())
    assert wins[0] + wins[1] + draws == quota
    assert mean_turns == pytest.approx(float(turns.mean()))
    assert nsims == orc.stats()["sims"]


def test_arena_agent_under_the_reference_server_loop():
    """Arena.play_games' loop body with host tensors and real queues around ArenaAgent."""
    import torch.multiprocessing as mp
    from azb200.arena import ArenaAgent
    B, sims, quota, seed = 12, 7, 20, 5
    nets = [FakeNN(4 * 6 * 7, 7, seed=90), FakeNN(4 * 6 * 7, 7, seed=91)]
    players = [_Model(n) for n in nets]
    ready, batch_q, resq = mp.Queue(), queue.Queue(), mp.Queue()
    ev, stop, pause = mp.Event(), mp.Event(), mp.Event()
    completed, played = mp.Value("i", 0), mp.Value("i", 0)
    pt, vt = torch.zeros(B, 7), torch.zeros(B, 3)
    ag = ArenaAgent(2, _C4Game, ready, ev, [[], []], pt, vt, batch_q, resq, completed, played, stop, pause,
                    _args(sims, quota), _is_arena=True, rng="philox", seed=seed)
    p2i = list(ag.player_to_index)
    ag.start()
    results = []
    while completed.value != 1:
        try:
            i = ready.get(timeout=1)
        except queue.Empty:
            continue
        assert i == 2
        policy, value = [], []
        data = batch_q.get()
        for player in range(len(players)):                       # Arena.pyx:266-271
            batch = data[player]
            if not isinstance(batch, list):
                p, v = players[player].process(batch)
                policy.append(p.to(pt.device)); value.append(v.to(vt.device))
        n = sum(len(p) for p in policy)
        pt[:n].copy_(torch.cat(policy)); vt[:n].copy_(torch.cat(value))
        ev.set()
    while True:
        try:
            results.append(resq.get(timeout=0.3))
        except queue.Empty:
            break
    ag.join(timeout=30)
    assert played.value == quota
    orc = _oracle_run(B, sims, quota, seed, p2i, nets, base=2 * B)
    _, turns, win = orc.results()
    assert [int(r[0].turns) for r in results] == turns.tolist()
    assert np.array_equal(np.stack([r[1] for r in results]), win)


def test_fused_two_model_evaluation_equals_the_generic_path(monkeypatch):
    """play_games with two real networks: the device-resident path (per-model row lists from select, compact tcgen05
    evaluation, no host synchronisation per simulation) against gather / NNetWrapper.process / scatter."""
    import azb200.arena as az_arena
    from azb200 import nnet as aznet
    B, sims, quota = 96, 10, 150
    nets = []
    for s in (1, 2):
        torch.manual_seed(s)
        nets.append(aznet.NNetWrapper(nnet=aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS).cuda().eval(), cuda=True,
                                      fused=True))
    out = []
    for mode in ("graph", "eager", "generic"):      # one CUDA graph per move-round / launch by launch / PyTorch gather-scatter
        if mode == "generic":
            monkeypatch.setattr(az_arena, "_fused_evaluators", lambda *a, **k: None)
        eng = az_arena.arena_engine(_C4Game, _args(sims, quota), B, rng="philox", seed=11)
        out.append(az_arena.play_games(eng, nets, (1, 0), sims=sims, round_graph=(mode == "graph")))
        eng.close()
    assert out[0] == out[1] == out[2]
    wins, draws, _, nsims = out[0]
    assert sum(wins) + draws == quota and nsims > quota * sims
