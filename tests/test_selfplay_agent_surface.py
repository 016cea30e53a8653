"""The reference-facing surface on the GPU: azb200.selfplay.SelfPlayAgent built
with the reference's constructor arguments and served by the loop body of
Coach.processSelfPlayBatches (Coach.py:337-342) with host tensors and real
multiprocessing queues/events -- output_queue / result_queue / games_played must
equal what the oracle produces for the same streams."""
import queue

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import _orc
from _fakenn import FakeNN

pytestmark = pytest.mark.gpu


class _Args(dict):
    __getattr__ = dict.__getitem__


class _C4Game:
    __module__ = "alphazero.envs.connect4.connect4"

    @staticmethod
    def max_turns():
        return 42


def _drain(q):
    out = []
    while True:
        try:
            out.append(q.get(timeout=0.2))
        except queue.Empty:
            return out


def test_coach_loop_drives_the_engine_agent():
    from azb200.selfplay import SelfPlayAgent
    B, sims, quota, seed = 24, 12, 30, 5
    args = _Args(gamesPerIteration=quota, numMCTSSims=sims, numFastSims=sims, numWarmupSims=sims, probFastSim=0.0,
                 cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1, add_root_noise=False,
                 add_root_temp=True, symmetricSamples=True, mctsResetThreshold=None, startTemp=1)
    ready, fileq, resq = mp.Queue(), mp.Queue(), mp.Queue()
    ev, stop, pause = mp.Event(), mp.Event(), mp.Event()
    completed, played = mp.Value("i", 0), mp.Value("i", 0)
    bt, pt, vt = torch.zeros(B, 4, 6, 7), torch.zeros(B, 7), torch.zeros(B, 3)
    ag = SelfPlayAgent(3, _C4Game, ready, ev, bt, pt, vt, fileq, resq, completed, played, stop, pause, args,
                       rng="philox", seed=seed)
    nn = FakeNN(4 * 6 * 7, 7, seed=17)
    ag.start()
    samples, results = [], []
    while completed.value != 1:                      # Coach.processSelfPlayBatches
        try:
            i = ready.get(timeout=1)
        except queue.Empty:
            continue
        assert i == 3
        p, v = nn(bt.numpy())
        pt.copy_(torch.from_numpy(p))
        vt.copy_(torch.from_numpy(v))
        ev.set()
        samples += _drain(fileq) if fileq.qsize() > 2000 else []
    samples += _drain(fileq)
    results += _drain(resq)
    ag.join(timeout=30)
    assert played.value == quota

    # the same streams through the oracle
    temps = _orc.temp_table(_orc.default_temp_scaling, 1, 42)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, B, rng_mode=_orc.RNG_PHILOX, seed=seed, game_id_base=3 * B,
                           add_root_temp=True, games_per_iteration=quota, temps=temps)
    while orc.stats()["games_played"] < quota:
        for _ in range(sims):
            obs = orc.generateBatch()
            orc.processBatch(*nn(obs))
        orc.playMoves(False)
    o_obs, o_pi, o_z, _ = orc.samples()
    assert len(samples) == len(o_obs) > 0
    assert np.array_equal(np.stack([s[0] for s in samples]), o_obs)
    assert np.array_equal(np.stack([s[1] for s in samples]), o_pi)
    assert np.array_equal(np.stack([s[2] for s in samples]), o_z)
    r_slot, r_turns, r_win = orc.results()
    assert [int(r[0].turns) for r in results] == r_turns.tolist()
    assert np.array_equal(np.stack([r[1] for r in results]), r_win)
    assert all(r[2] == 3 for r in results)
