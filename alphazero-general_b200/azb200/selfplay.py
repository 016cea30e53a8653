"""Host-side mirror of the reference's self-play interface on top of the engine.

``SelfPlayAgent`` keeps the constructor and the generateBatch / processBatch /
playMoves / run methods of alphazero/SelfPlayAgent.pyx:13-202, so
Coach.generateSelfPlayAgents / processSelfPlayBatches (Coach.py:291-361) drive
it unchanged: observation batches still travel through ``batch_tensor``, the
network's answers through ``policy_tensor`` / ``value_tensor`` (host tensors
-> host<->device copies every simulation), samples through ``output_queue``
and finished games through ``result_queue``.  It is a thread instead of a
process: the B games of the agent live on the GPU.

``DeviceSelfPlay`` is the B200-first driver: the NN reads and writes the
engine's device buffers in place (CUDA-graphed LeafEvaluator), tree kernels
and network run on separate streams ordered by CUDA events, optionally with
the games split into two cohorts so the tree work of one overlaps the
inference of the other.
"""
import threading
import time

import numpy as np
import torch

from .engine import SelfPlayEngine, default_temp_scaling, temp_table


def game_name(game_cls):
    """DeviceRules registry: which CUDA rule set implements this Game plugin."""
    name = getattr(game_cls, "AZB_GAME", None)
    if name:
        return name
    mod = getattr(game_cls, "__module__", "") or ""
    for c in getattr(game_cls, "__mro__", (game_cls,)):          # a subclass of a known plugin keeps its rules
        m = getattr(c, "__module__", "") or ""
        if "connect4" in m:
            return "connect4"
        if "brandubh" in m:
            return "brandubh"
        if "hnefatafl" in m and m.endswith("fastafl"):           # alphazero/envs/hnefatafl/fastafl.pyx, the 11x11 game
            return "hnefatafl"
    raise NotImplementedError(f"no device rules for game plugin {game_cls!r} (module {mod})")


def _max_turns(game_cls):
    fn = getattr(game_cls, "max_turns", None)      # brandubh's plugin lacks it -> GameState default None
    return fn() if fn else None


def engine_kwargs_from_args(game_cls, args, num_games, **over):
    """args (Coach.DEFAULT_ARGS keys) -> SelfPlayEngine keyword arguments."""
    g = lambda k, d=None: (args[k] if k in args else d)
    fn = g("temp_scaling_fn", default_temp_scaling)
    sims = max(int(g("numMCTSSims", 100) or 0), int(g("numFastSims", 0) or 0), int(g("numWarmupSims", 0) or 0), 1)
    kw = dict(
        game=game_name(game_cls), num_games=num_games,
        cpuct=g("cpuct", 1.25), fpu_reduction=g("fpu_reduction", 0.2),
        root_noise_frac=g("root_noise_frac", 0.1), root_policy_temp=g("root_policy_temp", 1.1),
        add_root_noise=g("add_root_noise", True), add_root_temp=g("add_root_temp", True),
        symmetric_samples=g("symmetricSamples", True), mcts_reset_threshold=g("mctsResetThreshold", None),
        games_per_iteration=g("gamesPerIteration", 0), max_sims_per_move=sims,
        temps=temp_table(fn, g("startTemp", 1), _max_turns(game_cls)),
    )
    kw.update(over)
    return kw


class FinalState:
    """What result_queue consumers read from the final game state
    (alphazero/utils.py:43-44 uses .turns)."""

    def __init__(self, turns, winstate):
        self.turns = int(turns)
        self._turns = int(turns)
        self.winstate = winstate

    def win_state(self):
        return self.winstate


class SelfPlayAgent(threading.Thread):
    def __init__(self, id, game_cls, ready_queue, batch_ready, batch_tensor, policy_tensor, value_tensor,
                 output_queue, result_queue, complete_count, games_played, stop_event, pause_event, args,
                 _is_arena=False, _is_warmup=False, engine=None, device=0, rng="philox", seed=None,
                 stream_ordered=False, step_graphs=None):
        super().__init__(daemon=True)
        if _is_arena:
            raise NotImplementedError("arena mode: use azb200.arena.ArenaAgent (same constructor; Arena.pyx:253-259)")
        self.id = id
        self.game_cls = game_cls
        self.ready_queue, self.batch_ready = ready_queue, batch_ready
        self.batch_tensor, self.policy_tensor, self.value_tensor = batch_tensor, policy_tensor, value_tensor
        self.batch_size = batch_tensor.shape[0]
        self.output_queue, self.result_queue = output_queue, result_queue
        self.complete_count, self.games_played = complete_count, games_played
        self.stop_event, self.pause_event = stop_event, pause_event
        self.args = args
        self._is_arena, self._is_warmup = _is_arena, _is_warmup
        self.fast = False
        self._rs = np.random.RandomState(seed)       # the worker's own stream: the fast-move coin
        if engine is None:
            engine = SelfPlayEngine(**engine_kwargs_from_args(
                game_cls, args, self.batch_size, device=device, rng=rng,
                seed=int(self._rs.randint(0, 2 ** 31 - 1)) if seed is None else seed,
                game_id_base=id * self.batch_size))
        self.engine = engine
        if _is_warmup:
            # SelfPlayAgent.pyx:48-52: policy = 1/A, value = 1/(P+1) (float32)
            self.engine.policy.copy_(torch.full((engine.A,), 1 / engine.A).expand(engine.B, engine.A))
            self.engine.value.copy_(torch.full((3,), 1 / 3).expand(engine.B, 3))
        # stream_ordered: the host tensors stay the transport (every batch still crosses PCIe both ways), but their
        # validity is ordered by CUDA events instead of host synchronisation: generateBatch returns once the copy into
        # batch_tensor is ENQUEUED and records `batch_event`; the NN server makes its stream wait on that event before
        # it uploads batch_tensor, and records `answer_event` (set by the server through `set_answer_event`) after it
        # has enqueued the copies into policy_tensor / value_tensor.  Requires pinned host tensors.
        self.stream_ordered = stream_ordered
        # step_graphs (stream-ordered mode only): the per-simulation launch sequences of this agent -- upload of the
        # answers, expand/backup, select, download of the next observation batch -- are captured once as CUDA graphs
        # and replayed, which removes most of the per-simulation host work
        self.step_graphs = stream_ordered if step_graphs is None else (step_graphs and stream_ordered)
        self._g_first = self._g_mid = self._g_last = None
        self._rounds = 0
        self.batch_event = None
        self.answer_event = None
        self._counted = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.batches = 0            # generateBatch calls (simulation batches) so far

    def _check_pause(self):
        while self.pause_event.is_set():
            time.sleep(.1)

    def run(self):
        # the agent's kernels and copies go to its own stream, so several agents and the
        # NN server overlap on one GPU
        self.stream = torch.cuda.Stream(device=self.engine.obs.device)
        with torch.cuda.stream(self.stream):
            self._run()

    def _run(self):
        try:
            while not self.stop_event.is_set() and self.games_played.value < self.args.gamesPerIteration:
                self._check_pause()
                self.fast = self._rs.random_sample() < self.args.probFastSim
                sims = self.args.numFastSims if self.fast else self.args.numMCTSSims \
                    if not self._is_warmup else self.args.numWarmupSims
                if self._is_warmup:
                    self.engine.warmup_sims(sims)
                elif self.step_graphs and self._rounds > 0:      # the first round runs eagerly (one-time kernel setup)
                    self._graphed_round(sims)
                else:
                    for _ in range(sims):
                        if self.stop_event.is_set(): break
                        self.generateBatch()
                        if self.stop_event.is_set(): break
                        self.processBatch()
                if self.stop_event.is_set(): break
                self.playMoves()
                self._rounds += 1
            with self.complete_count.get_lock():
                self.complete_count.value += 1
            # the reference closes its process-local end of output_queue here; a thread shares the
            # queue object with the consumer, so it is left open
        except Exception:
            import traceback
            print(traceback.format_exc())

    # ---- stream-ordered protocol with captured step graphs -------------------------------------------
    def _capture(self, fn):
        from .nnet import capture_graph
        return capture_graph(fn, self.stream)

    def _download_batch(self):
        from .nnet import download
        download(self.batch_tensor, self.engine.obs)

    def _upload_answers(self):
        from .nnet import upload
        upload(self.engine.policy, self.policy_tensor)
        upload(self.engine.value, self.value_tensor)

    def _publish_batch(self):
        ev = torch.cuda.Event()
        ev.record()
        self.batch_event = ev
        self.d2h_bytes += self.batch_tensor.numel() * 4
        self.batches += 1
        self.ready_queue.put(self.id)

    def _await_answers(self):
        self.batch_ready.wait()
        self.batch_ready.clear()
        if self.stop_event.is_set():
            return False
        if self.answer_event is not None:
            torch.cuda.current_stream().wait_event(self.answer_event)
        self.h2d_bytes += (self.policy_tensor.numel() + self.value_tensor.numel()) * 4
        return True

    def _graphed_round(self, sims):
        """generateBatch / processBatch of one move-round, `sims` times, as three replayed graphs:
        [select, obs -> batch_tensor], [answers -> engine, expand/backup, select, obs -> batch_tensor], [answers -> engine,
        expand/backup]."""
        eng = self.engine
        if self._g_first is None:
            self._prepare_graphs()
        self._g_first.replay()
        self._publish_batch()
        for _ in range(sims - 1):
            if not self._await_answers():
                return
            self._g_mid.replay()
            self._publish_batch()
        if not self._await_answers():
            return
        self._g_last.replay()

    def round_with_server(self, server, sims):
        """One move-round (`sims` x generateBatch -> NN server -> processBatch) as ONE replayed CUDA graph on this
        agent's stream, the body of the NN server (Coach.py:337-342: process(batch_tensor), copies into policy_tensor /
        value_tensor) captured inside it.  Every batch still travels through the pinned HOST tensors in both directions
        -- observations device -> batch_tensor -> device, answers device -> policy / value tensors -> device -- but the
        host issues one launch per agent and round instead of ~5 per agent and simulation: with the per-step graphs the
        issuing Python thread is busy 100 % of the time and bounds the protocol (profiles/r2_e2e_scaling.md).
        playMoves stays with the caller."""
        key = int(sims)
        graphs = self.__dict__.setdefault("_round_graphs", {})
        g = graphs.get(key)
        if g is None:
            eng = self.engine

            def whole_round():
                eng.select(stream=self.stream)
                self._download_batch()
                for s_ in range(key):
                    server._body(self)                              # upload batch_tensor, network, answers -> host tensors
                    self._upload_answers()
                    eng.expand_backup(stream=self.stream)
                    if s_ + 1 < key:
                        eng.select(stream=self.stream)
                        self._download_batch()
            with torch.cuda.stream(self.stream):
                if not graphs.get("warm"):
                    server._body(self)                              # lazy one-time setup of the evaluator: not capturable
                    graphs["warm"] = True
                    self.stream.synchronize()
                g = graphs[key] = self._capture(whole_round)
        with torch.cuda.stream(self.stream):
            g.replay()
        self.batches += key
        self.d2h_bytes += key * self.batch_tensor.numel() * 4
        self.h2d_bytes += key * (self.policy_tensor.numel() + self.value_tensor.numel()) * 4

    def _prepare_graphs(self):
        eng = self.engine

        def first():
            eng.select(stream=self.stream)
            self._download_batch()

        def mid():
            self._upload_answers()
            eng.expand_backup(stream=self.stream)
            eng.select(stream=self.stream)
            self._download_batch()

        def last():
            self._upload_answers()
            eng.expand_backup(stream=self.stream)
        self._g_first, self._g_mid, self._g_last = self._capture(first), self._capture(mid), self._capture(last)

    def generateBatch(self):
        self._check_pause()
        self.engine.select()
        if self._is_warmup:
            return
        if self.stream_ordered:
            self._download_batch()
            ev = torch.cuda.Event()
            ev.record()
            self.batch_event = ev
        else:
            self.batch_tensor.copy_(self.engine.obs)             # device -> caller's (host) tensor
        if not self.batch_tensor.is_cuda:
            self.d2h_bytes += self.batch_tensor.numel() * 4
        self.batches += 1
        self.ready_queue.put(self.id)

    def processBatch(self):
        if self._is_warmup:
            self.engine.expand_backup()           # engine policy/value rows hold the warmup constants
            return
        self.batch_ready.wait()
        self.batch_ready.clear()
        if self.stop_event.is_set():
            return
        if self.stream_ordered and self.answer_event is not None:
            torch.cuda.current_stream().wait_event(self.answer_event)
        self._upload_answers()
        if not self.policy_tensor.is_cuda:
            self.h2d_bytes += (self.policy_tensor.numel() + self.value_tensor.numel()) * 4
        self.engine.expand_backup()

    def playMoves(self):
        self._check_pause()
        self.engine.play_moves(self.fast)
        self.engine.check_errors()
        slot, turns, win = self.engine.drain_results()
        for i in range(len(slot)):
            self.result_queue.put((FinalState(turns[i], win[i]), win[i].copy(), self.id))
        if len(slot):
            obs, pi, z, _ = self.engine.drain_samples()
            self.d2h_bytes += obs.nbytes + pi.nbytes + z.nbytes
            if hasattr(self.output_queue, "put_block"):
                # azb200.coach.ExampleQueue: the same examples in the same order as one block instead of one queue
                # item each (a move-round of 8192 games ends ~400 games = ~16 k examples; the per-item loop costs the
                # host longer than the GPU needs for the next 30 simulations)
                self.output_queue.put_block(torch.from_numpy(obs), torch.from_numpy(pi), torch.from_numpy(z))
            else:
                for i in range(len(obs)):
                    self.output_queue.put((obs[i], pi[i], z[i]))
            played = self.engine.games_played()
            new = played - self._counted
            self._counted = played
            if new:
                with self.games_played.get_lock():
                    self.games_played.value += new


class DeviceSelfPlay:
    """Device-resident self-play: select -> CUDA-graphed ResNet -> expand/backup
    per simulation, tree kernels on ``tree_stream`` and the network on
    ``nn_stream``; with cohorts == 2 the two halves of the games alternate so
    that tree work of one half overlaps inference of the other."""

    def __init__(self, engine, nnet, cohorts=1, precision=None, use_graph=True, channels_last=False, fused=None,
                 split=None, round_graph=None, skip_terminal=True, dedup=None):
        """precision: operand precision of the leaf evaluator, see azb200.nn_tc.make_evaluator -- "bf16x2" (default:
        hand-written tcgen05 kernels, within 1e-5 of the reference's fp32 module), "fp16", "bf16" (opt-in performance
        modes), "fp32" / "tf32" / "cudnn-bf16" (PyTorch / cuDNN).  fused: None = the hand-written kernels whenever they
        cover the network and the precision; False = cuDNN; "tc-r1" / "mma" = the round-1 bf16-only kernels.
        dedup: leaf de-duplication (engine.set_leaf_dedup: games whose leaves have the same observation share one
        evaluation; bit-identical results) -- None = on whenever the evaluator takes the engine's compact row list."""
        assert cohorts in (1, 2)
        from . import nn_tc
        self.engine = engine
        self.cohorts = cohorts
        B = engine.B
        if fused is False and (precision is None or precision in nn_tc.PRECISIONS):
            precision = {None: "tf32", "bf16": "cudnn-bf16"}.get(precision, "tf32")
        precision = precision or nn_tc.default_precision(nnet)
        kernel = fused if fused in ("tc-r1", "mma") else None
        hand = kernel is not None or (precision in nn_tc.PRECISIONS and nn_tc.supported(nnet))
        if cohorts == 2 and split is None:
            split = B // 2
        bounds = [0, B] if cohorts == 1 else [0, split, B]
        self.ranges = [(bounds[i], bounds[i + 1] - bounds[i]) for i in range(cohorts)]
        self.compact = bool(hand and kernel != "mma" and skip_terminal and cohorts == 1)
        self.dedup = bool(self.compact and not getattr(engine, "arena", False) and (True if dedup is None else dedup))
        if self.dedup or getattr(engine, "leaf_dedup", False):
            engine.set_leaf_dedup(self.dedup)
        if self.compact:
            # compact evaluation: only the leaves that need the network (the reference evaluates terminal leaves
            # too and discards the answers, SelfPlayAgent.pyx:116-123 / MCTS.pyx:234-235)
            self.evals = [nn_tc.make_evaluator(nnet, engine.obs, engine.policy, engine.value, precision=precision, kernel=kernel,
                                               rows=engine.nn_rows, count=engine.nn_count_ptr, max_batch=c)
                          for f, c in self.ranges]
        else:
            self.evals = [nn_tc.make_evaluator(nnet, engine.obs[f:f + c], engine.policy[f:f + c], engine.value[f:f + c],
                                               precision=precision, kernel=kernel, use_graph=use_graph,
                                               channels_last=channels_last)
                          for f, c in self.ranges]
        self.precision = getattr(self.evals[0], "precision", precision)
        fused = hand
        dev = engine.obs.device
        self.tree_stream = torch.cuda.Stream(device=dev)
        self.nn_stream = torch.cuda.Stream(device=dev)
        self.ev_sel = [torch.cuda.Event() for _ in self.ranges]
        self.ev_nn = [torch.cuda.Event() for _ in self.ranges]
        self.launches = 0
        # A whole move-round (sims x (select, network, expand/backup) + playMoves) as ONE CUDA graph: every launch
        # argument is constant and all state lives in device memory, so the graph replays unchanged; it removes the
        # ~20 us of launch / cross-stream event latency per simulation.  Fused evaluators only (the cuDNN evaluator
        # is itself a graph), single cohort (the kernels of a cohort form one dependency chain anyway).
        self.round_graph = (bool(fused) and cohorts == 1) if round_graph is None else round_graph
        self._graphs = {}
        self._eager_rounds = 0

    def run_round(self, sims, fast=False):
        """One move-round: ``sims`` simulations for every game, then playMoves.
        Asynchronous: returns once everything is enqueued on tree_stream."""
        eng, T, N = self.engine, self.tree_stream, self.nn_stream
        cur = torch.cuda.current_stream()
        if self.round_graph:
            key = (sims, bool(fast))
            g = self._graphs.get(key)
            if g is None and self._eager_rounds >= 1:      # capture after one eager round (lazy one-time setup done)
                # capture_graph, not the torch.cuda.graph context manager (which runs gc.collect() and
                # torch.cuda.empty_cache() first): the kernels are launched through the C ABI, nothing allocates
                from .nnet import capture_graph

                def whole_round():
                    (f, c), = self.ranges
                    eng.select(f, c, stream=T)
                    for s in range(sims):
                        self.evals[0](stream=T)
                        if s + 1 < sims:
                            eng.expand_backup_select(f, c, stream=T)      # processBatch(s) + generateBatch(s+1)
                    eng.expand_backup(f, c, stream=T)
                    eng.play_moves(fast, stream=T)
                T.wait_stream(cur)
                g = capture_graph(whole_round, T)
                self._graphs[key] = g
            if g is not None:
                T.wait_stream(cur)
                with torch.cuda.stream(T):
                    g.replay()
                cur.wait_stream(T)
                # kernels of this repository inside the graph: select, sims x (evaluator + expand/backup[+select]), playMoves
                # (play, finalize, emit, emit_reset)
                self.launches += sims * (getattr(self.evals[0], "kernels_per_call", 1) + 1) + 5
                return
            self._eager_rounds += 1
        T.wait_stream(cur)
        for s in range(sims):
            for h, (f, c) in enumerate(self.ranges):
                if s > 0:
                    T.wait_event(self.ev_nn[h])
                    eng.expand_backup(f, c, stream=T)
                eng.select(f, c, stream=T)
                self.ev_sel[h].record(T)
                N.wait_event(self.ev_sel[h])
                self.evals[h](stream=N)
                self.ev_nn[h].record(N)
        for h, (f, c) in enumerate(self.ranges):
            T.wait_event(self.ev_nn[h])
            eng.expand_backup(f, c, stream=T)
        eng.play_moves(fast, stream=T)
        cur.wait_stream(T)
        self.launches += sims * len(self.ranges) * (getattr(self.evals[0], "kernels_per_call", 1) + 2) + 4

    def run_round_warmup(self, sims, fast=False):
        """Tree-only round (numWarmupSims): one fused kernel + playMoves."""
        self.engine.warmup_sims(sims)
        self.engine.play_moves(fast)
        self.launches += 5
