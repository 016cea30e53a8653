"""run_selfplay_iteration (the Coach self-play phase on the engine): quota,
fast-move coin, sample files in the reference's three-file format, and equality
with the oracle driven by the same coin sequence."""
import os

import numpy as np
import pytest
import torch

import _orc
from _fakenn import warmup_outputs

pytestmark = pytest.mark.gpu


class _C4Game:
    __module__ = "alphazero.envs.connect4.connect4"
    @staticmethod
    def max_turns(): return 42
    @staticmethod
    def action_size(): return 7
    @staticmethod
    def observation_size(): return (4, 6, 7)
    @staticmethod
    def num_players(): return 2
    @staticmethod
    def has_draw(): return True


def test_warmup_iteration_matches_oracle_and_writes_reference_files(tmp_path):
    from azb200.coach import run_selfplay_iteration, save_iteration_samples
    args = dict(process_batch_size=64, gamesPerIteration=100, numWarmupSims=12, probFastSim=0.6, numFastSims=4,
                numMCTSSims=12, add_root_noise=False, add_root_temp=True, symmetricSamples=True)
    res = run_selfplay_iteration(_C4Game, None, args, seed=9, warmup=True)
    wins, draws, avg_len = res.game_results()
    assert len(res.result_turns) >= 100 and wins[0] + wins[1] + draws == len(res.result_turns) and 7 <= avg_len <= 42
    base = save_iteration_samples(res, str(tmp_path), "run", 3)
    d, p, v = (torch.load(base + s, weights_only=False) for s in ("-data.pkl", "-policy.pkl", "-value.pkl"))
    assert os.path.basename(base) == "iteration-0003"
    assert d.shape[1:] == (4, 6, 7) and p.shape == (d.shape[0], 7) and v.shape == (d.shape[0], 3) and d.dtype == torch.float32
    # the oracle with the same coin sequence and streams
    temps = _orc.temp_table(_orc.default_temp_scaling, 1, 42)
    orc = _orc.OracleAgent(_orc.GAME_CONNECT4, 64, rng_mode=_orc.RNG_PHILOX, seed=9, add_root_temp=True,
                           games_per_iteration=100, temps=temps)
    rs = np.random.RandomState(9)
    while orc.stats()["games_played"] < 100:
        fast = bool(rs.random_sample() < 0.6)
        for _ in range(4 if fast else 12):       # numFastSims if fast else numWarmupSims (SelfPlayAgent.pyx:85-86)
            orc.generateBatch(); orc.processBatch(*warmup_outputs(64, 7))
        orc.playMoves(fast)
    o, pi, z, _ = orc.samples()
    assert np.array_equal(d.numpy(), o) and np.array_equal(p.numpy(), pi) and np.array_equal(v.numpy(), z)
    assert np.array_equal(res.result_turns, orc.results()[1])


def test_nn_iteration_with_fused_and_cudnn_evaluators():
    from azb200 import nnet as aznet
    from azb200.coach import run_selfplay_iteration
    torch.manual_seed(0)
    model = aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS).cuda().eval()
    args = dict(process_batch_size=128, gamesPerIteration=128, numMCTSSims=16, numFastSims=4, probFastSim=0.5,
                add_root_noise=True, add_root_temp=True, symmetricSamples=True)
    for fused in (True, False):
        res = run_selfplay_iteration(_C4Game, model, args, seed=1, fused=fused)
        assert len(res.result_turns) >= 128 and res.data.shape[0] == res.policy.shape[0] == res.value.shape[0] > 0
        assert torch.allclose(res.policy.sum(1), torch.ones(res.policy.shape[0]), atol=1e-5)
        assert torch.all(res.value.sum(1) == 1)
        assert res.sims > 0


def test_device_resident_examples_equal_the_host_drain():
    """device_samples=True keeps the examples on the GPU (chunked device-to-device drain): same tensors, same order."""
    from azb200.coach import run_selfplay_iteration
    import azb200.coach as coach
    args = dict(process_batch_size=64, gamesPerIteration=150, numWarmupSims=6, probFastSim=0.3, symmetricSamples=True)
    host = run_selfplay_iteration(_C4Game, None, args, seed=4, warmup=True)
    orig = coach._DeviceSampleSink.__init__
    try:
        coach._DeviceSampleSink.__init__ = lambda self, engine, chunk=0: orig(self, engine, chunk=700)   # several chunks
        dev = run_selfplay_iteration(_C4Game, None, args, seed=4, warmup=True, device_samples=True)
    finally:
        coach._DeviceSampleSink.__init__ = orig
    assert dev.data.is_cuda and dev.data.shape[0] == host.data.shape[0] > 1400
    for a, b in ((dev.data, host.data), (dev.policy, host.policy), (dev.value, host.value)):
        assert torch.equal(a.cpu(), b)
    assert np.array_equal(dev.result_turns, host.result_turns) and dev.sims == host.sims


class _RefShapedCoach:
    """The slice of the reference Coach the mixin plugs into: its fields (Coach.py:176-207) and the consumer side of
    saveIterationSamples / get_game_results (Coach.py:366-376, utils.py:34-54), restated for the test."""

    def __init__(self, game_cls, args, tmp):
        import torch.multiprocessing as mp
        self.game_cls, self.args = game_cls, args
        self.args.data, self.args.run_name = str(tmp), "run"
        self.warmup, self.sample_time = True, 0
        import types
        self.train_net = self.self_play_net = types.SimpleNamespace(nnet=None)       # warmup iterations never call it
        self.stop_train = mp.Event()
        self.file_queue, self.result_queue = mp.Queue(), mp.Queue()
        self.completed, self.games_played = mp.Value("i", 0), mp.Value("i", 0)

    def saveIterationSamples(self, iteration):              # the reference's loop: one queue item per example
        n = self.file_queue.qsize()
        data = torch.zeros([n, *self.game_cls.observation_size()])
        pol, val = torch.zeros([n, self.game_cls.action_size()]), torch.zeros([n, 3])
        for i in range(n):
            d, p, v = self.file_queue.get()
            data[i], pol[i], val[i] = torch.from_numpy(d), torch.from_numpy(p), torch.from_numpy(v)
        return data, pol, val


def test_mixin_fills_the_coach_fields_queues_and_files(tmp_path):
    """class MyCoach(GpuSelfPlayMixin, Coach): the three self-play phase methods + saveIterationSamples."""
    from azb200.coach import ExampleQueue, GpuSelfPlayMixin

    class Args(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    class MyCoach(GpuSelfPlayMixin, _RefShapedCoach):
        pass

    args = Args(process_batch_size=48, gamesPerIteration=100, numWarmupSims=5, probFastSim=0.4, symmetricSamples=True,
                model_gating=True, workers=3)
    c = MyCoach(_C4Game, args, tmp_path)
    np.random.seed(3)
    c.generateSelfPlayAgents()
    c.processSelfPlayBatches(1)
    assert c.games_played.value == 100 and c.completed.value == 3 and c.sample_time > 0
    n = c.file_queue.qsize()
    assert isinstance(c.file_queue, ExampleQueue) and n > 100 * 7
    first = c.file_queue.get()                              # the per-example queue protocol still works ...
    assert first[0].shape == (4, 6, 7) and first[1].shape == (7,) and first[2].shape == (3,) and c.file_queue.qsize() == n - 1
    rest = _RefShapedCoach.saveIterationSamples(c, 1)       # ... for the reference's own consumer loop
    assert rest[0].shape[0] == n - 1 and c.file_queue.empty()
    results = [c.result_queue.get(timeout=10) for _ in range(100)]
    assert all(r[1].sum() == 1 and 7 <= r[0].turns <= 42 and np.array_equal(r[0].win_state(), r[1]) for r in results)
    # second iteration: the mixin's saveIterationSamples writes the reference's three files from the tensors
    c.killSelfPlayAgents()
    c.generateSelfPlayAgents()
    c.processSelfPlayBatches(2)
    n2 = c.file_queue.qsize()
    c.saveIterationSamples(2)
    base = tmp_path / "run" / "iteration-0002"
    d, p, v = (torch.load(str(base) + s, weights_only=False) for s in ("-data.pkl", "-policy.pkl", "-value.pkl"))
    assert d.shape == (n2, 4, 6, 7) and p.shape == (n2, 7) and v.shape == (n2, 3) and c.file_queue.empty()
    assert torch.allclose(p.sum(1), torch.ones(n2), atol=1e-5) and torch.all(v.sum(1) == 1)
