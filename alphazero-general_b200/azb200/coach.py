"""Self-play phase of alphazero/Coach.py on the engine.

``GpuSelfPlayMixin`` overrides the three self-play phase methods of the
reference Coach (generateSelfPlayAgents / processSelfPlayBatches /
killSelfPlayAgents, Coach.py:291-361,401-435) so that
``class MyCoach(GpuSelfPlayMixin, Coach)`` keeps the rest of the loop
(saveIterationSamples, processGameResults, train, gating, GUI fields) unchanged:
samples and results still arrive through ``file_queue`` / ``result_queue`` and
``games_played`` / ``completed`` / ``sample_time`` are maintained.

``run_selfplay_iteration`` is the same phase without the reference package: it
plays ``gamesPerIteration`` games on one GPU and returns (and optionally saves,
in the reference's three-file format, Coach.py:364-386) the training examples.
"""
import os
import pickle
import time

import numpy as np
import torch

from .engine import SelfPlayEngine
from .selfplay import DeviceSelfPlay, FinalState, engine_kwargs_from_args


def get_iter_file(iteration):
    """alphazero/utils.py:15-16"""
    return f"iteration-{iteration:04d}.pkl"


class SelfPlayResult:
    def __init__(self, obs, pi, z, slot_r, turns_r, win_r, sims, seconds):
        self.data, self.policy, self.value = obs, pi, z
        self.result_slots, self.result_turns, self.result_winstates = slot_r, turns_r, win_r
        self.sims, self.seconds = sims, seconds
        self.last_round_seconds = np.zeros(0)        # host time per move-round (diagnostic, set by run_selfplay_iteration)

    def game_results(self, num_players=2):
        """alphazero/utils.py:34-54 get_game_results"""
        wins = [int(self.result_winstates[:, p].sum()) for p in range(num_players)]
        draws = int(self.result_winstates[:, num_players].sum())
        n = len(self.result_turns)
        return wins, draws, (float(self.result_turns.sum()) / n if n else 0)


class _DeviceSampleSink:
    """Examples stay on the GPU: drained device-to-device from the engine's sample ring into chunks allocated ahead of
    the rounds that fill them (no allocation, no host copy of the examples inside the loop)."""

    def __init__(self, engine, chunk=1 << 18):
        self.eng, self.chunk = engine, int(chunk)
        self.dev = engine.obs.device
        self.chunks, self.fill = [], 0
        self._grow(self.chunk)

    def _grow(self, rows):
        self.chunks.append([torch.empty((rows,) + self.eng.obs_shape, device=self.dev), torch.empty(rows, self.eng.A, device=self.dev),
                            torch.empty(rows, 3, device=self.dev), 0])

    def drain(self, n):
        o, p, z, k = self.chunks[-1]
        if k + n > o.shape[0]:
            self._grow(max(self.chunk, n))
            o, p, z, k = self.chunks[-1]
        got = self.eng.drain_samples_into(o[k:k + n], p[k:k + n], z[k:k + n])
        self.chunks[-1][3] = k + got

    def tensors(self):
        return tuple(torch.cat([c[i][:c[3]] for c in self.chunks]) for i in range(3))

    def to_host(self):
        """-> pinned host tensors (one copy per chunk and tensor, then one synchronisation)."""
        n = sum(c[3] for c in self.chunks)
        out = tuple(torch.empty((n,) + tuple(self.chunks[0][i].shape[1:]), pin_memory=True) for i in range(3))
        k = 0
        for c in self.chunks:
            for i in range(3):
                out[i][k:k + c[3]].copy_(c[i][:c[3]], non_blocking=True)
            k += c[3]
        torch.cuda.synchronize(self.dev)
        return out


def run_selfplay_iteration(game_cls, nnet_module, args, device=0, seed=0, warmup=False, engine=None, fused=None,
                           precision=None, game_id_base=0, stop_event=None, progress=None, device_samples=False):
    """One self-play phase (SelfPlayAgent.run for a single GPU-resident agent):
    until gamesPerIteration games are counted, draw the fast-move coin, run
    numFastSims / numMCTSSims (numWarmupSims in a warmup iteration) simulations
    for every game, play the moves.  Returns a SelfPlayResult (host tensors; CUDA
    tensors that never left the device with ``device_samples``)."""
    g = lambda k, d=None: (args[k] if k in args else d)
    # the reference runs `workers` agents of process_batch_size games each at once (Coach.py:294-323): one engine holds
    # them all (never more slots than games the iteration asks for)
    B = int(g("process_batch_size", 256)) * max(1, int(g("workers", 1) or 1))
    if g("gamesPerIteration"):
        B = max(1, min(B, int(g("gamesPerIteration"))))
    own_engine = engine is None
    if own_engine:
        engine = SelfPlayEngine(**engine_kwargs_from_args(game_cls, args, B, device=device, rng="philox", seed=seed,
                                                          game_id_base=game_id_base))
    else:
        engine.reset_games(seed)
        engine.set_quota(g("gamesPerIteration", 0))
    drv = None
    if not warmup:
        # leaf evaluator: the hand-written tcgen05 kernels at their default precision ("bf16x2": within 1e-5 of the
        # reference's fp32 module) wherever they cover the network, else cuDNN TF32 (the reference's own arithmetic);
        # a narrower type only when the caller asks for it (args.nn_precision / precision=)
        # args.leaf_dedup (optional, default on): games whose leaves have the same observation share one evaluation
        drv = DeviceSelfPlay(engine, nnet_module, cohorts=1, precision=precision or g("nn_precision"), channels_last=True,
                             fused=fused, dedup=g("leaf_dedup", None))
    rs = np.random.RandomState(seed)
    quota = int(g("gamesPerIteration"))
    rslot, rturns, rwin = [], [], []
    # the examples are drained device-to-device inside the loop either way; for host results they cross PCIe once at
    # the end, into pinned memory (torch's caching host allocator keeps the block for the next iteration)
    sink = _DeviceSampleSink(engine)
    t0 = time.time()
    round_times = []
    while engine.games_played() < quota and not (stop_event is not None and stop_event.is_set()):
        round_times.append(time.time())
        fast = bool(rs.random_sample() < g("probFastSim", 0.0))
        # sims = numFastSims if fast else (numMCTSSims if not warmup else numWarmupSims)      (SelfPlayAgent.pyx:85-86)
        if warmup:
            engine.warmup_sims(int(g("numFastSims", 20)) if fast else int(g("numWarmupSims", 5)))
            engine.play_moves(fast)
        else:
            drv.run_round(int(g("numFastSims", 20)) if fast else int(g("numMCTSSims", 100)), fast)
        engine.check_errors()
        n = engine.sample_count()
        if n > 0:
            sink.drain(n)
        s, t, w = engine.drain_results()
        if len(s):
            rslot.append(s); rturns.append(t); rwin.append(w)
            if progress is not None:
                progress(engine.games_played())
    cat = lambda xs, shape, dt: np.concatenate(xs) if xs else np.zeros(shape, dt)
    A, obs_shape, nsims = engine.A, engine.obs_shape, engine.stats()["sims"]
    dt = time.time() - t0
    round_times.append(time.time())
    round_seconds = np.diff(np.asarray(round_times))                            # diagnostic: host time per move-round
    if own_engine:
        torch.cuda.synchronize(engine.obs.device)
        engine.close()                                   # the node pool goes back to the driver now, not at some later GC
    tensors = sink.tensors() if device_samples else sink.to_host()
    res = SelfPlayResult(*tensors, cat(rslot, (0,), np.int32), cat(rturns, (0,), np.int32), cat(rwin, (0, 3), np.uint8),
                         nsims, dt)
    res.last_round_seconds = round_seconds
    return res


def save_iteration_samples(result, data_dir, run_name, iteration):
    """Coach.saveIterationSamples (Coach.py:364-386): three torch.save files."""
    folder = os.path.join(data_dir, run_name)
    os.makedirs(folder, exist_ok=True)
    base = os.path.join(folder, get_iter_file(iteration).replace(".pkl", ""))
    torch.save(result.data, base + "-data.pkl", pickle_protocol=pickle.HIGHEST_PROTOCOL)
    torch.save(result.policy, base + "-policy.pkl", pickle_protocol=pickle.HIGHEST_PROTOCOL)
    torch.save(result.value, base + "-value.pkl", pickle_protocol=pickle.HIGHEST_PROTOCOL)
    return base


class ExampleQueue:
    """``file_queue`` of the reference Coach (an mp.Queue of (obs, pi, z) numpy triples, SelfPlayAgent.pyx:194-196) backed
    by the iteration's example tensors: ``qsize / empty / get / put`` behave like the queue for the reference's
    saveIterationSamples loop (Coach.py:366-376), but three tensors are stored instead of one pickled item per example
    (an iteration of 8192 Connect4 games is ~400 k examples)."""

    def __init__(self):
        self.blocks, self.row, self.singles = [], 0, []

    def put_block(self, data, policy, value):
        if data.shape[0]:
            self.blocks.append((data, policy, value))

    def put(self, item, *a, **k):
        self.singles.append(item)

    def qsize(self):
        return sum(int(b[0].shape[0]) for b in self.blocks) - self.row + len(self.singles)

    def empty(self):
        return self.qsize() == 0

    def get(self, *a, **k):
        if self.blocks:
            d, p, v = self.blocks[0]
            i = self.row
            self.row += 1
            if self.row == d.shape[0]:
                self.blocks.pop(0)
                self.row = 0
            return d[i].numpy(), p[i].numpy(), v[i].numpy()
        if self.singles:
            return self.singles.pop(0)
        import queue
        raise queue.Empty

    def take_tensors(self):
        """Everything queued, as three tensors (None if single items are queued too); empties the queue."""
        if self.singles or not self.blocks:
            return None
        blocks, self.blocks = self.blocks, []
        row, self.row = self.row, 0
        blocks[0] = tuple(t[row:] for t in blocks[0])
        return tuple(torch.cat([b[i] for b in blocks]) for i in range(3))


def game_with_defaults(game_cls):
    """A Game plugin that does not derive from GameState (brandubh's fastafl.Game) lacks the defaults of
    alphazero/Game.py:55-63 that Coach.__init__ (Coach.py:161) and SelfPlayAgent.playMoves (SelfPlayAgent.pyx:157)
    call: max_turns() -> None, has_draw() -> True.  Returns the class itself or a subclass that supplies them (the
    device rules follow the plugin's module, so the subclass keeps them)."""
    missing = {k: staticmethod(v) for k, v in (("max_turns", lambda: None), ("has_draw", lambda: True))
               if not hasattr(game_cls, k)}
    if not missing:
        return game_cls
    return type(game_cls.__name__, (game_cls,), dict(missing, __module__=game_cls.__module__))


class GpuSelfPlayMixin:
    """Mix into the reference Coach: ``class GpuCoach(GpuSelfPlayMixin, Coach): pass``.
    A Game plugin without device rules (anything but Connect4 / brandubh today) keeps the reference's own
    SelfPlayAgent processes: every override below then defers to the Coach it is mixed into."""

    def _has_device_rules(self):
        from .selfplay import game_name
        try:
            game_name(self.game_cls)
            return True
        except NotImplementedError:
            return False

    def generateSelfPlayAgents(self):
        if not self._has_device_rules():
            return super().generateSelfPlayAgents()
        self._gpu_engine_args = dict(seed=int(np.random.randint(0, 2 ** 31 - 1)))

    def processSelfPlayBatches(self, iteration):
        if not self._has_device_rules():
            return super().processSelfPlayBatches(iteration)
        nnet = self.self_play_net if self.args.model_gating else self.train_net
        t0 = time.time()

        def progress(n):
            self.games_played.value = int(n)
            self.sample_time = (time.time() - t0) / max(n, 1)

        res = run_selfplay_iteration(self.game_cls, nnet.nnet, self.args, warmup=self.warmup,
                                     stop_event=self.stop_train, progress=progress, **self._gpu_engine_args)
        for i in range(len(res.result_turns)):
            w = res.result_winstates[i]
            self.result_queue.put((FinalState(res.result_turns[i], w), w, 0))
        if not isinstance(self.file_queue, ExampleQueue):
            self.file_queue = ExampleQueue()
        self.file_queue.put_block(res.data, res.policy, res.value)
        self.games_played.value = min(int(self.args.gamesPerIteration), len(res.result_turns))
        self.completed.value = self.args.workers
        self.sample_time = res.seconds / max(self.games_played.value, 1)
        if hasattr(self, "writer"):
            self.writer.add_scalar("loss/sample_time", self.sample_time, iteration)

    def saveIterationSamples(self, iteration):
        """Coach.saveIterationSamples (Coach.py:364-386): the same three files, written from the tensors."""
        t = self.file_queue.take_tensors() if isinstance(self.file_queue, ExampleQueue) else None
        if t is None:
            return super().saveIterationSamples(iteration)
        print(f'Saving {t[0].shape[0]} samples')
        try:                                                    # the GUI polls coach.state (Coach.py:128-137)
            from alphazero.Coach import TrainState
            self.state = TrainState.SAVE_SAMPLES
        except ImportError:
            TrainState = None
        save_iteration_samples(SelfPlayResult(*t, None, None, None, 0, 0.0), self.args.data, self.args.run_name, iteration)
        if TrainState is not None:
            self.state = TrainState.STANDBY

    def killSelfPlayAgents(self):
        if not self._has_device_rules():
            return super().killSelfPlayAgents()
        import torch.multiprocessing as mp
        self.agents = []
        self.file_queue, self.result_queue = ExampleQueue(), mp.Queue()
        self.completed, self.games_played = mp.Value("i", 0), mp.Value("i", 0)
