"""BASELINE config 5 on one process per GPU: the iteration loop of alphazero/Coach.py (learn, Coach.py:225-289) with every
phase on the device -- self-play on the engine (coach.run_selfplay_iteration), the sample window on the GPU
(samples.SampleWindow), training through the loop body of NNetWrapper.train (NNetWrapper.py:122-190) over that
window, the comparison with the self-play model as a batched arena on the engine (arena.play_games) and the
reference's gating rule (Coach.compareToPast, Coach.py:528-570).

With torch.distributed initialised (one rank per GPU) self-play shards by game and the examples are gathered to rank
0 (azb200.distributed), which trains and gates and broadcasts the accepted weights (``ddp_train=True``: every rank
trains on its slice of each batch, gradients summed over NCCL, azb200.distributed.train_steps_sharded); without it
everything runs on one GPU.  The reference's GUI fields, tensorboard writer and checkpoint files are not part of this loop.
"""
import time

import numpy as np
import torch

from . import nnet as aznet
from .arena import arena_engine, play_games
from .coach import run_selfplay_iteration
from .engine import SelfPlayEngine
from .selfplay import engine_kwargs_from_args
from .samples import SampleWindow, loss_pi, loss_v

DEFAULTS = dict(
    numIters=2, numWarmupIters=1, gamesPerIteration=256, process_batch_size=256, numMCTSSims=100, numFastSims=20,
    numWarmupSims=5, probFastSim=0.75, train_batch_size=1024, train_steps_per_iteration=64, autoTrainSteps=True,
    averageTrainSteps=False, minTrainHistoryWindow=4, maxTrainHistoryWindow=20, trainHistoryIncrementIters=2,
    lr=1e-2, momentum=0.9, weight_decay=1e-4, value_loss_weight=1.5, compareWithPast=True, pastCompareFreq=1,
    arenaCompare=128, arenaTemp=0.25, model_gating=True, max_gating_iters=None, min_next_model_winrate=0.52,
    use_draws_for_winrate=True, cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1,
    add_root_noise=True, add_root_temp=True, symmetricSamples=True, mctsResetThreshold=None, startTemp=1,
    ddp_train=False, graph_train=False)


class _A(dict):
    __getattr__ = dict.__getitem__


class _GraphedStep:
    """One training step (forward, both losses, backward, SGD update) on static input tensors, captured as a CUDA
    graph: the DEFAULT_ARGS net is ~200 small kernels per step, i.e. launch-bound when issued one by one."""

    def __init__(self, net, optimizer, boards, pis, vs, value_loss_weight):
        self.boards, self.pis, self.vs = (torch.empty_like(t) for t in (boards, pis, vs))
        self.losses = torch.zeros(2, device=boards.device)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            out_pi, out_v = net(self.boards)
            l_pi, l_v = loss_pi(self.pis, out_pi), loss_v(self.vs, out_v, value_loss_weight)
            (l_pi + l_v).backward()
            optimizer.step()
            self.losses.copy_(torch.stack((l_pi.detach(), l_v.detach())))

    def __call__(self, boards, pis, vs):
        self.boards.copy_(boards); self.pis.copy_(pis); self.vs.copy_(vs)
        self.graph.replay()
        return self.losses


def _eager_step(net, optimizer, boards, target_pis, target_vs, value_loss_weight):
    # a function of its own: when it returns nothing references the step's autograd graph any more, so the
    # parameters' AccumulateGrad nodes (bound to the stream of this step) are gone before a capture builds its own
    out_pi, out_v = net(boards)
    l_pi, l_v = loss_pi(target_pis, out_pi), loss_v(target_vs, out_v, value_loss_weight)
    optimizer.zero_grad()
    (l_pi + l_v).backward()
    optimizer.step()
    return torch.stack((l_pi.detach(), l_v.detach()))


def train_steps(wrapper, optimizer, loader, steps, value_loss_weight, graph_after=3):
    """The loop body of NNetWrapper.train (NNetWrapper.py:131-165): -> (mean policy loss, mean value loss).
    After `graph_after` eager steps on full batches (cuDNN has chosen its algorithms, the momentum buffers exist) the
    step is captured once per (wrapper, batch shape) and replayed; ragged last batches of an epoch run eagerly.
    graph_after=None: every step eager."""
    net = wrapper.nnet
    net.train()
    acc = None                                        # device-side sum of batch-weighted (l_pi, l_v): no sync per step
    n = 0
    step = full_eager = 0
    cache = wrapper.__dict__.setdefault("_graphed_steps", {})
    while step < steps:
        for boards, target_pis, target_vs in loader:
            if step == steps:
                break
            step += 1
            b = boards.size(0)
            key = (tuple(boards.shape), id(optimizer))
            g = cache.get(key)
            if g is None and graph_after is not None and boards.is_cuda and b == loader.batch_size and full_eager >= graph_after:
                g = cache[key] = _GraphedStep(net, optimizer, boards, target_pis, target_vs, value_loss_weight)
            if g is not None:
                losses = g(boards, target_pis, target_vs)
            else:
                losses = _eager_step(net, optimizer, boards, target_pis, target_vs, value_loss_weight)
                full_eager += int(b == loader.batch_size)
            acc = losses.double() * b if acc is None else acc + losses.double() * b
            n += b
    net.eval()
    if not n:
        return 0.0, 0.0
    lp, lv = (acc / n).tolist()
    return lp, lv


def winrate_of_first(wins, draws, use_draws):
    """Arena.__update_winrates / PlayerStats.update (Arena.pyx:124-130): draws count half for everybody."""
    games = sum(wins) + (draws if use_draws else 0)
    return ((wins[0] + (draws if use_draws else 0) / len(wins)) / games) if games else 0.0


class GpuCoach:
    def __init__(self, game_cls, args=None, device=0, seed=0, net_args=None):
        self.game_cls = game_cls
        self.args = _A(DEFAULTS)
        self.args.update(args or {})
        self.device, self.seed = device, seed
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        torch.manual_seed(seed)
        dev = torch.device("cuda", device)
        mk = lambda: aznet.NNetWrapper(nnet=aznet.ResNet(tuple(game_cls.observation_size()), game_cls.action_size(), 3,
                                                         **(net_args or aznet.DEFAULT_NET_ARGS)).to(dev), cuda=True, fused=True)
        self.train_net, self.self_play_net = mk(), mk()
        self.self_play_net.nnet.load_state_dict(self.train_net.nnet.state_dict())
        self.optimizer = torch.optim.SGD(self.train_net.nnet.parameters(), lr=self.args.lr, momentum=self.args.momentum,
                                         weight_decay=self.args.weight_decay)
        self.window = SampleWindow(device=dev)
        self.self_play_iter, self.gating_counter = 0, 0
        self.history = []
        self._sp_engine = None

    def _fresh(self, wrapper):
        wrapper._fused_eval = {}             # folded weights are snapshots: rebuild after the weights changed

    def learn(self):
        a = self.args
        for it in range(1, a.numIters + 1):
            rec = dict(iteration=it)
            warmup = it <= a.numWarmupIters or self.self_play_iter == 0          # Coach.py:232-238
            t0 = time.time()
            from .distributed import shard_games
            sp_args = dict(a, gamesPerIteration=shard_games(int(a.gamesPerIteration), self.rank, self.world)[1])
            if self._sp_engine is None:                                          # one node pool for all iterations
                self._sp_engine = SelfPlayEngine(**engine_kwargs_from_args(
                    self.game_cls, sp_args, int(a.process_batch_size), device=self.device, rng="philox", seed=self.seed,
                    game_id_base=self.rank * a.process_batch_size))
            res = run_selfplay_iteration(self.game_cls, self.self_play_net.nnet, sp_args,
                                         device=self.device, seed=self.seed + 1000 * it + self.rank, warmup=warmup,
                                         engine=self._sp_engine, device_samples=True)
            obs, pi, z = res.data, res.policy, res.value                        # CUDA tensors, never copied to the host
            local_windows = self.world > 1 and a.ddp_train == "local"           # every rank keeps its own examples
            if self.world > 1 and not local_windows:
                from .distributed import gather_examples_to_rank0
                t_g = time.time()
                obs, pi, z = gather_examples_to_rank0(obs, pi, z)
                torch.cuda.synchronize()
                rec["gather_seconds"] = time.time() - t_g
            if self.world > 1:                          # every rank must be in the same mode (the gating decision is broadcast)
                flags = [None] * self.world
                torch.distributed.all_gather_object(flags, bool(warmup))
                rec["warmup_by_rank"] = flags
                tot = torch.tensor([float(res.sims)], device=torch.device("cuda", self.device), dtype=torch.float64)
                torch.distributed.all_reduce(tot)
                rec["sims_all_ranks"] = float(tot.item())
            rec.update(selfplay_seconds=time.time() - t0, selfplay_loop_seconds=res.seconds, sims=res.sims, warmup=warmup,
                       selfplay_rounds=int(len(res.last_round_seconds)), slowest_round_seconds=float(np.max(res.last_round_seconds, initial=0.0)),
                       samples=int(obs.shape[0]) if self.rank == 0 else 0, game_results=res.game_results())
            loader, steps, used = None, 0, []
            if self.rank == 0 or local_windows:
                self.window.add_iteration(it, obs, pi, z)
                self.window.evict_before(it - a.maxTrainHistoryWindow)
                if local_windows:
                    used = self.window.window(it, a)
                else:
                    loader, used = self.window.loader(it, a)
                    steps = self.window.train_steps(used, a)
            t0 = time.time()
            if local_windows:
                from .distributed import train_steps_local_windows
                lp, lv, steps, gsamples = train_steps_local_windows(self.train_net.nnet, self.optimizer, self.window, it, a,
                                                                    a.value_loss_weight, torch.device("cuda", self.device))
                losses = (lp, lv)
                rec["samples_all_ranks"] = gsamples
                self._fresh(self.train_net)
            elif self.world > 1 and a.ddp_train:
                # every rank trains: rank 0 draws the batches from its window, the rows of each batch are split
                # over the ranks, gradients summed (azb200.distributed.train_steps_sharded)
                from .distributed import train_steps_sharded
                losses = train_steps_sharded(self.train_net.nnet, self.optimizer, loader, steps, a.value_loss_weight,
                                             torch.device("cuda", self.device))
                self._fresh(self.train_net)
            elif self.rank == 0:
                losses = train_steps(self.train_net, self.optimizer, loader, steps, a.value_loss_weight,
                                     graph_after=3 if a.graph_train else None)
                self._fresh(self.train_net)
            if self.rank == 0:
                rec["loss_pi"], rec["loss_v"] = losses
                rec.update(train_steps=steps, train_seconds=time.time() - t0, window=used)
                if a.compareWithPast and (it - 1) % a.pastCompareFreq == 0:
                    rec.update(self.compare_to_past(it))
            if not a.model_gating and self.rank == 0:
                # Coach.processSelfPlayBatches plays with train_net when gating is off (Coach.py:338): keep the network
                # self-play uses in step with it after every training phase
                self.self_play_net.nnet.load_state_dict(self.train_net.nnet.state_dict())
                self._fresh(self.self_play_net)
            if self.world > 1:                                                # everybody self-plays with rank 0's decision
                for p in list(self.self_play_net.nnet.parameters()) + list(self.self_play_net.nnet.buffers()):
                    torch.distributed.broadcast(p.data, 0)
                # ... and knows it: self_play_iter decides warm-up mode on every rank (Coach.py:232-238)
                t = torch.tensor([self.self_play_iter, self.gating_counter], device=torch.device("cuda", self.device),
                                 dtype=torch.int64)
                torch.distributed.broadcast(t, 0)
                self.self_play_iter, self.gating_counter = int(t[0].item()), int(t[1].item())
                self._fresh(self.self_play_net)
            self.history.append(rec)
        return self.history

    def compare_to_past(self, model_iter):
        """Coach.compareToPast (Coach.py:528-570): new net (model 0) against the self-play net, gating."""
        a = self.args
        # The reference's Arena starts `workers` agents, each with its own shuffle of the sides (SelfPlayAgent.pyx:44-46),
        # so the new network moves first in some games and second in others.  One engine = one side assignment: play the
        # match as two halves with opposite assignments (which half gets the extra game of an odd match is drawn).
        rs = np.random.RandomState(self.seed + model_iter)
        first = [0, 1]
        rs.shuffle(first)
        n_a = (a.arenaCompare + 1) // 2
        halves = [(tuple(first), n_a), (tuple(first[::-1]), a.arenaCompare - n_a)]
        t0 = time.time()
        wins, draws, sims, turns_sum, t_setup, rounds, slowest = [0, 0], 0, 0, 0.0, 0.0, 0, 0.0
        for k, (p2i, games) in enumerate(halves):
            if games <= 0:
                continue
            eng = arena_engine(self.game_cls, dict(a, gamesPerIteration=games), min(games, 4096), device=self.device,
                               rng="philox", seed=self.seed + 7 * model_iter + k)
            w, d, mean_turns, ns = play_games(eng, [self.train_net, self.self_play_net], p2i, sims=a.numMCTSSims,
                                              fast_sims=a.numFastSims, prob_fast=a.probFastSim, coin=rs)
            eng.close()
            wins = [wins[0] + w[0], wins[1] + w[1]]
            draws, sims, turns_sum = draws + d, sims + ns, turns_sum + mean_turns * games
            t_setup += play_games.setup_seconds
            rounds += int(len(play_games.last_round_seconds))
            slowest = max(slowest, float(np.max(play_games.last_round_seconds, initial=0.0)))
        t_play = time.time() - t0
        winrate = winrate_of_first(wins, draws, a.use_draws_for_winrate)
        out = dict(arena_wins=wins, arena_draws=draws, arena_winrate=winrate, arena_seconds=time.time() - t0, arena_play_seconds=t_play, arena_setup_seconds=t_setup, arena_sims=sims,
                   arena_rounds=rounds, arena_slowest_round_seconds=slowest, arena_sides=[h[0] for h in halves])
        if a.model_gating and winrate < a.min_next_model_winrate and (a.max_gating_iters is None or self.gating_counter < a.max_gating_iters):
            self.gating_counter += 1
            out["accepted"] = False
        else:
            self.self_play_iter = model_iter
            self.self_play_net.nnet.load_state_dict(self.train_net.nnet.state_dict())
            self._fresh(self.self_play_net)
            self.gating_counter = 0
            out["accepted"] = True
        return out
