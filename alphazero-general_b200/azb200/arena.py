"""Batched arena on the engine: alphazero/Arena.pyx:208-328 (Arena.play_games, batched branch) and the
``_is_arena`` mode of alphazero/SelfPlayAgent.pyx (:23,44-47,62-73,104-132,144-150,167-168).

The engine runs in arena mode (``azb_config.arena``): game i lives in slots 2i / 2i+1, the search trees of env player
0 / 1; per simulation only the tree of the player to move searches, after every move both trees follow it.  Two
surfaces on top:

``ArenaAgent``   the reference-shaped agent: same constructor as ``SelfPlayAgent(..., _is_arena=True)``; per simulation
                 it puts the per-model observation batches on ``output_queue`` (host tensors), the caller answers in
                 ``policy_tensor`` / ``value_tensor`` in concatenated model order -- the loop body of Arena.play_games
                 (Arena.pyx:262-275) drives it unchanged; results arrive on ``result_queue``.
``play_games``   device-resident: the models are evaluated on the GPU rows in place (gather by model, evaluate,
                 scatter); returns wins per model and draws like Arena.play_games.

One deliberate difference from the reference: ``processBatch`` there reads row ``batch_indices[i]`` for game i
(SelfPlayAgent.pyx:144), i.e. the batch *position -> game* list used as a *game -> position* map, which is only right
while every game of the worker has the same player to move.  Here game i always receives the answer computed for its
own observation (equal to the reference whenever the reference is self-consistent, e.g. one game per agent, which is
how the parity tests pin it).
"""
import threading
import time

import numpy as np
import torch

from .engine import SelfPlayEngine
from .selfplay import FinalState, engine_kwargs_from_args


def arena_engine(game_cls, args, num_games, **over):
    """SelfPlayEngine in arena mode for `num_games` games (2 x num_games slots)."""
    kw = engine_kwargs_from_args(game_cls, args, 2 * num_games, **over)
    kw.update(arena=True, add_root_noise=False, add_root_temp=False,
              temps=np.full(1, float(args["arenaTemp"] if "arenaTemp" in args else 0.25), dtype=np.float64))
    kw["max_sims_per_move"] = max(int(args["numMCTSSims"] if "numMCTSSims" in args else 100),
                                  int(args["numFastSims"] if "numFastSims" in args else 0), 1)
    return SelfPlayEngine(**kw)


class _Round:
    """One simulation round of an arena engine: which rows search, and which model evaluates them."""

    def __init__(self, engine, player_to_index):
        self.engine = engine
        self.p2i = torch.as_tensor(list(player_to_index), dtype=torch.int64, device=engine.obs.device)

    def rows_by_model(self):
        """-> list over models of int64 row (slot) tensors, ascending by game (batch_indices of SelfPlayAgent.pyx:118)."""
        pl = self.engine.arena_players().to(torch.int64)
        rows = torch.nonzero(pl >= 0).flatten()
        model = self.p2i[pl[rows]]
        return [rows[model == m] for m in range(len(self.p2i))]


def play_games(engine, models, player_to_index=(0, 1), sims=None, stop_event=None, progress=None, round_graph=True,
               fast_sims=None, prob_fast=0.0, coin=None):
    """Arena.play_games on the device: `models[m].process(batch) -> (pi, v)` (NNetWrapper surface), model
    `player_to_index[p]` moves for env player p.  Plays until the engine's games_per_iteration quota is reached.
    prob_fast / fast_sims: the arena-mode agent of the reference draws the fast-move coin per move-round as well
    (SelfPlayAgent.pyx:84-86: numFastSims simulations when it comes up); coin = RandomState it is drawn from.
    -> (wins per model index, draws, mean turns, simulations run)."""
    sims_full = int(sims or 100)
    fast_sims = int(fast_sims or sims_full)
    coin = coin or np.random.RandomState(0)
    rnd = _Round(engine, player_to_index)
    wins, draws, turns = [0] * len(models), 0, []
    quota = engine.quota
    t0 = time.time()
    # fast path: both models are NNetWrappers whose network the tcgen05 evaluator covers -- every simulation is
    # select -> model 0 on its row list -> model 1 on its row list -> expand/backup, all on the device, no host sync
    evals = _fused_evaluators(engine, models, player_to_index)
    play_games.setup_seconds = time.time() - t0
    # ... and then a whole move-round (sims x (select, model 0, model 1, expand/backup) + playMoves) is ONE CUDA graph:
    # every launch argument is constant (row lists and counters live on the device), so after one eager round the
    # graph is captured and replayed per move -- same kernels, same order, same results (tests/test_arena.py)
    graphs, rounds = {}, 0
    use_graph = round_graph and evals is not None
    gstream = torch.cuda.Stream(device=engine.obs.device) if use_graph else None

    side = torch.cuda.Stream(device=engine.obs.device) if evals is not None else None
    e_fork, e_join = torch.cuda.Event(), torch.cuda.Event()

    def fused_round(sims, stream=None):
        stream = stream or torch.cuda.current_stream()
        engine.select(stream=stream)
        for s in range(sims):
            # the two models read disjoint row lists and write disjoint rows: evaluate them side by side (a fork /
            # join in the captured graph) -- at gating sizes each evaluation is a handful of CTAs
            e_fork.record(stream)
            side.wait_event(e_fork)
            evals[1](stream=side)
            e_join.record(side)
            evals[0](stream=stream)
            stream.wait_event(e_join)
            if s + 1 < sims:
                engine.expand_backup_select(stream=stream)          # processBatch(s) + generateBatch(s+1), one launch
        engine.expand_backup(stream=stream)
        engine.play_moves(False, stream=stream)

    stamps = []
    while engine.games_played() < quota and not (stop_event is not None and stop_event.is_set()):
        stamps.append(time.time())
        sims = fast_sims if (prob_fast > 0.0 and coin.random_sample() < prob_fast) else sims_full
        graph = graphs.get(sims)
        if evals is not None:
            if use_graph and graph is None and rounds >= 1:
                # not the torch.cuda.graph context manager: it runs gc.collect() and torch.cuda.empty_cache() first
                # (hundreds of ms right after a training phase); nothing here allocates through torch
                from .nnet import capture_graph
                gstream.wait_stream(torch.cuda.current_stream())
                graph = graphs[sims] = capture_graph(lambda: fused_round(sims, gstream), gstream)
            if graph is not None:
                cur = torch.cuda.current_stream()
                gstream.wait_stream(cur)
                with torch.cuda.stream(gstream):
                    graph.replay()
                cur.wait_stream(gstream)
            else:
                fused_round(sims)
            rounds += 1
        else:
            for _ in range(sims):
                engine.select()
                for m, rows in enumerate(rnd.rows_by_model()):
                    if rows.numel() == 0:
                        continue
                    pi, v = models[m].process(engine.obs.index_select(0, rows))
                    engine.policy.index_copy_(0, rows, pi.to(engine.policy.dtype))
                    engine.value.index_copy_(0, rows, v.to(engine.value.dtype))
                engine.expand_backup()
            engine.play_moves(False)
        engine.check_errors()
        slot, t, win = engine.drain_results()
        for i in range(len(slot)):
            if len(turns) >= quota:
                break                                           # results beyond the quota are not scored (utils.py:34-54)
            for p in range(len(player_to_index)):
                wins[player_to_index[p]] += int(win[i][p])
            draws += int(win[i][len(player_to_index)])
            turns.append(int(t[i]))
        if progress is not None and len(slot):
            progress(len(turns), time.time() - t0)
    stamps.append(time.time())
    play_games.last_round_seconds = np.diff(np.asarray(stamps))                 # diagnostic: host time per move-round
    return wins, draws, (float(np.mean(turns)) if turns else 0.0), engine.stats()["sims"]


def _fused_evaluators(engine, models, player_to_index):
    """One compact tcgen05 evaluator per model over the engine's per-model row lists, or None."""
    if len(models) != 2 or not all(getattr(m, "fused", False) and hasattr(m, "nnet") for m in models):
        return None
    from .nn_tc import TensorCoreEvaluator, supported
    if not all(supported(m.nnet) for m in models):
        return None
    engine.arena_set_player_to_index(player_to_index)
    return [TensorCoreEvaluator(m.nnet, engine.obs, engine.policy, engine.value, precision=getattr(m, "precision", None),
                                rows=engine.arena_rows(k), count=(lambda k=k: engine.arena_count_ptr(k)),
                                max_batch=engine.B // 2)
            for k, m in enumerate(models)]


class ArenaAgent(threading.Thread):
    """SelfPlayAgent(..., _is_arena=True) (SelfPlayAgent.pyx:14-16): ``batch_tensor`` is the reference's per-player
    list (unused), ``policy_tensor`` / ``value_tensor`` hold the answers of one simulation round in concatenated
    model order, ``output_queue`` carries the observation batches, ``result_queue`` the finished games."""

    def __init__(self, id, game_cls, ready_queue, batch_ready, batch_tensor, policy_tensor, value_tensor,
                 output_queue, result_queue, complete_count, games_played, stop_event, pause_event, args,
                 _is_arena=True, _is_warmup=False, engine=None, device=0, rng="philox", seed=None):
        super().__init__(daemon=True)
        assert _is_arena and not _is_warmup
        self.id, self.game_cls, self.args = id, game_cls, args
        self.ready_queue, self.batch_ready = ready_queue, batch_ready
        self.policy_tensor, self.value_tensor = policy_tensor, value_tensor
        self.batch_size = policy_tensor.shape[0]                # SelfPlayAgent.pyx:23-24
        self.output_queue, self.result_queue = output_queue, result_queue
        self.complete_count, self.games_played = complete_count, games_played
        self.stop_event, self.pause_event = stop_event, pause_event
        self._rs = np.random.RandomState(seed)
        self.player_to_index = list(range(game_cls.num_players() if hasattr(game_cls, "num_players") else 2))
        self._rs.shuffle(self.player_to_index)                  # SelfPlayAgent.pyx:44-46
        if engine is None:
            engine = arena_engine(game_cls, args, self.batch_size, device=device, rng=rng,
                                  seed=int(self._rs.randint(0, 2 ** 31 - 1)) if seed is None else seed,
                                  game_id_base=id * self.batch_size)
        self.engine = engine
        self.batch_indices = None
        self._rows = None
        self._counted = 0

    def run(self):
        try:
            with torch.cuda.stream(torch.cuda.Stream(device=self.engine.obs.device)):
                while not self.stop_event.is_set() and self.games_played.value < self.args.gamesPerIteration:
                    # the arena-mode agent draws the fast-move coin as well (SelfPlayAgent.pyx:84-86)
                    fast = self._coin() < (self.args["probFastSim"] if "probFastSim" in self.args else 0.0)
                    for _ in range(self.args.numFastSims if fast else self.args.numMCTSSims):
                        if self.stop_event.is_set(): break
                        self.generateBatch()
                        if self.stop_event.is_set(): break
                        self.processBatch()
                    if self.stop_event.is_set(): break
                    self.playMoves()
            with self.complete_count.get_lock():
                self.complete_count.value += 1
        except Exception:
            import traceback
            print(traceback.format_exc())

    def _coin(self):
        """np.random.random_sample() of SelfPlayAgent.run, from the agent's own stream (the games' streams live on the
        device, one per slot)."""
        return float(self._rs.random_sample())

    def generateBatch(self):
        while self.pause_event.is_set():
            time.sleep(.1)
        self.engine.select()
        rows = _Round(self.engine, self.player_to_index).rows_by_model()
        batch = []
        for r in rows:                                           # SelfPlayAgent.pyx:124-131: tensor per model, [] if none
            batch.append(self.engine.obs.index_select(0, r).cpu() if r.numel() else [])
        self._rows = torch.cat(rows)
        self.batch_indices = (self._rows // 2).tolist()          # games in batch order (model 0's first)
        self.output_queue.put(batch)
        self.ready_queue.put(self.id)

    def processBatch(self):
        self.batch_ready.wait()
        self.batch_ready.clear()
        if self.stop_event.is_set():
            return
        n = self._rows.numel()
        dev = self.engine.policy.device
        self.engine.policy.index_copy_(0, self._rows, self.policy_tensor[:n].to(dev))
        self.engine.value.index_copy_(0, self._rows, self.value_tensor[:n].to(dev))
        self.engine.expand_backup()

    def playMoves(self):
        self.engine.play_moves(False)
        self.engine.check_errors()
        slot, turns, win = self.engine.drain_results()
        for i in range(len(slot)):
            self.result_queue.put((FinalState(turns[i], win[i]), win[i].copy(), self.id))
        if len(slot):
            played = self.engine.games_played()
            new, self._counted = played - self._counted, played
            if new:
                with self.games_played.get_lock():
                    self.games_played.value += new
