"""Host-side logic that needs no GPU: temperature tables, args mapping, the
device-rules registry, the ResNet mirror against the reference's own module."""
import os
import sys

import numpy as np
import pytest
import torch

import _refdriver
from azb200 import default_temp_scaling, temp_table
from azb200 import nnet as aznet
from azb200.selfplay import engine_kwargs_from_args, game_name


def test_default_temperature_table_connect4():
    t = temp_table(default_temp_scaling, 1, 42)
    # halves every int(0.15*42)=6 plies, floor 0.2 (alphazero/utils.py:19-27)
    assert t[:5].tolist() == [1, 1, 1, 1, 1] and t[5] == 0.5 and t[11] == 0.25 and t[17] == 0.2 and t[41] == 0.2


def test_temperature_table_without_max_turns_is_constant():
    assert np.all(temp_table(default_temp_scaling, 1, None) == 1)


@pytest.mark.skipif(not _refdriver.available(), reason="oracle/_ref not built")
def test_temperature_table_matches_reference_schedule():
    _refdriver._import_ref()
    from alphazero.utils import default_temp_scaling as ref_fn
    cur = 1
    ours = temp_table(default_temp_scaling, 1, 42, n=60)
    for turn in range(60):
        cur = ref_fn(cur, turn, 42)
        assert ours[turn] == cur


def test_args_mapping_and_registry():
    class G:
        __module__ = "alphazero.envs.connect4.connect4"
        @staticmethod
        def max_turns(): return 42
    args = dict(cpuct=4, fpu_reduction=0.4, numMCTSSims=200, numFastSims=40, gamesPerIteration=64,
                symmetricSamples=False, add_root_noise=False)
    kw = engine_kwargs_from_args(G, args, 8)
    assert kw["game"] == "connect4" and kw["cpuct"] == 4 and kw["max_sims_per_move"] == 200
    assert kw["games_per_iteration"] == 64 and kw["symmetric_samples"] is False and kw["add_root_temp"] is True

    class B:
        __module__ = "alphazero.envs.brandubh.fastafl"
    assert game_name(B) == "brandubh"
    assert np.all(engine_kwargs_from_args(B, {}, 4)["temps"] == 1)      # plugin lacks max_turns -> None

    class X:
        __module__ = "alphazero.envs.othello.othello"
    with pytest.raises(NotImplementedError):
        game_name(X)


@pytest.mark.skipif(not _refdriver.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("preset", ["default", "connect4_train"])
def test_resnet_mirror_matches_reference_module(preset):
    """Same state_dict keys/shapes as alphazero/NNetArchitecture.ResNet and the same
    process() outputs (strict fp32 on CPU; north-star tolerance 1e-5)."""
    _refdriver._import_ref()
    from alphazero.NNetArchitecture import ResNet as RefResNet
    from alphazero.utils import dotdict
    game = _refdriver.game_class("connect4")
    na = aznet.DEFAULT_NET_ARGS if preset == "default" else aznet.CONNECT4_TRAIN_NET_ARGS
    torch.manual_seed(0)
    ref = RefResNet(game, dotdict(na)).eval()
    ours = aznet.ResNet.for_game(game, dict(na)).eval()
    sd = ref.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    # non-trivial BN statistics
    for k, v in sd.items():
        if k.endswith("running_var"):
            v.uniform_(0.5, 1.5)
        elif k.endswith("running_mean"):
            v.normal_(0, 0.2)
    ours.load_state_dict(sd)
    x = torch.rand(64, 4, 6, 7)
    with torch.no_grad():
        lp, lv = ref(x)
        want = torch.exp(lp), torch.exp(lv)
    got = aznet.NNetWrapper(nnet=ours, cuda=False).process(x)
    assert torch.allclose(got[0], want[0], atol=1e-5, rtol=0) and torch.allclose(got[1], want[1], atol=1e-5, rtol=0)
    assert torch.allclose(got[0].sum(1), torch.ones(64), atol=1e-5)


def test_example_queue_keeps_the_per_example_protocol():
    """azb200.coach.ExampleQueue: Coach.file_queue backed by tensors -- qsize / empty / get hand out the reference's
    (obs, pi, z) numpy triples in emission order across blocks and single items; take_tensors drains the blocks."""
    import queue
    import torch
    from azb200.coach import ExampleQueue
    q = ExampleQueue()
    assert q.empty() and q.qsize() == 0 and q.take_tensors() is None
    with pytest.raises(queue.Empty):
        q.get()
    mk = lambda n, base: (torch.arange(n * 4 * 6 * 7, dtype=torch.float32).view(n, 4, 6, 7) + base,
                          torch.full((n, 7), float(base)), torch.full((n, 3), float(base)))
    q.put_block(*mk(3, 100))
    q.put_block(*mk(0, 0))                                  # an empty block is not queued
    q.put_block(*mk(2, 200))
    assert q.qsize() == 5 and not q.empty()
    d, p, v = q.get()
    assert d.shape == (4, 6, 7) and d[0, 0, 0] == 100 and p.tolist() == [100.0] * 7 and v.tolist() == [100.0] * 3
    assert q.qsize() == 4
    t = q.take_tensors()                                     # the rest, still in order, the consumed row left out
    assert t[0].shape == (4, 4, 6, 7) and t[1][:, 0].tolist() == [100.0, 100.0, 200.0, 200.0] and q.empty()
    assert float(t[0][0, 0, 0, 0]) == 100 + 4 * 6 * 7
    q.put_block(*mk(1, 7))
    q.put((np.zeros((4, 6, 7), np.float32), np.ones(7, np.float32), np.ones(3, np.float32)))      # a reference-style item
    assert q.qsize() == 2 and q.take_tensors() is None      # mixed content: only the per-item path serves it
    assert q.get()[1][0] == 7 and q.get()[1][0] == 1 and q.empty()


def test_mixin_defers_to_the_reference_coach_for_games_without_device_rules():
    """GpuSelfPlayMixin: a plugin the engine has no rules for keeps the reference's own agents."""
    from azb200.coach import GpuSelfPlayMixin
    calls = []

    class Base:
        def generateSelfPlayAgents(self): calls.append("gen")
        def processSelfPlayBatches(self, iteration): calls.append(("proc", iteration))
        def killSelfPlayAgents(self): calls.append("kill")
        def saveIterationSamples(self, iteration): calls.append(("save", iteration))

    class Othello:
        __module__ = "alphazero.envs.othello.othello"

    class Coach(GpuSelfPlayMixin, Base):
        game_cls = Othello
        file_queue = object()

    c = Coach()
    c.generateSelfPlayAgents(); c.processSelfPlayBatches(3); c.saveIterationSamples(3); c.killSelfPlayAgents()
    assert calls == ["gen", ("proc", 3), ("save", 3), "kill"] and not hasattr(c, "_gpu_engine_args")
