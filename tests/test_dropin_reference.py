"""The drop-in claim against the REAL reference classes (compiled from /root/reference by oracle/build_ref.py into
oracle/_ref): alphazero.Coach.Coach, alphazero.Arena.Arena, alphazero.NNetWrapper.NNetWrapper and the real Game plugins.

  * class C(GpuSelfPlayMixin, Coach): C(Game, nnet, args).learn() -- the reference's own loop (Coach.py:225-288) runs a
    warm-up iteration and a network iteration with the self-play phase on the engine; saveIterationSamples writes the
    three .pkl files, the reference's Coach.train consumes them and writes checkpoints, compareToPast gates with the
    reference's own Arena.
  * Arena.play_games(use_batched_mcts=True) with the import swap of INTEGRATION.md (alphazero.Arena.SelfPlayAgent =
    azb200.arena.ArenaAgent): the reference's server loop drives the engine-backed agents; results are consistent with
    the engine's own records, and on fed seeds equal the device-resident driver's win counts.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.gpu


def _ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")      # Coach checkpoints hold args (classes, functions)
    import alphazero.Coach as coach_mod
    return coach_mod


def _args(coach_mod, tmp, **over):
    from alphazero.utils import dotdict
    base = dotdict({k: (dotdict(v) if isinstance(v, dict) else v) for k, v in coach_mod.DEFAULT_ARGS.items()})   # get_args mutates
    base.update(dict(
        run_name="dropin", cuda=True, workers=2, process_batch_size=64, gamesPerIteration=128, numIters=2, numWarmupIters=1,
        numMCTSSims=16, numFastSims=4, numWarmupSims=5, probFastSim=0.5, train_batch_size=64, autoTrainSteps=False,
        train_steps_per_iteration=3, compareWithBaseline=False, compareWithPast=True, arenaCompare=8, arena_batch_size=4,
        arenaBatched=True, min_next_model_winrate=0.0, checkpoint=os.path.join(str(tmp), "checkpoint"),
        data=os.path.join(str(tmp), "data"), startIter=0, load_model=True, _num_players=3))
    base.update(over)
    return base


@pytest.mark.parametrize("game", ["connect4", "brandubh", "hnefatafl"])
def test_real_coach_learns_with_the_engine_as_its_self_play_phase(game, tmp_path, monkeypatch):
    coach_mod = _ref()
    from alphazero.NNetWrapper import NNetWrapper
    from azb200.coach import ExampleQueue, GpuSelfPlayMixin, game_with_defaults
    monkeypatch.chdir(tmp_path)                     # SummaryWriter('runs/<run>')
    if game == "connect4":
        from alphazero.envs.connect4.connect4 import Game
        obs, A = (4, 6, 7), 7
    elif game == "brandubh":
        from alphazero.envs.brandubh.fastafl import Game
        Game = game_with_defaults(Game)             # the plugin lacks GameState's max_turns / has_draw (Coach.py:161)
        obs, A = (5, 7, 7), 588
    else:
        from alphazero.envs.hnefatafl.fastafl import Game
        Game = game_with_defaults(Game)
        obs, A = (5, 11, 11), 2420
    args = _args(coach_mod, tmp_path)
    games = 128
    if game == "hnefatafl":      # 11x11 games last hundreds of plies and an example is 12 KB: a small iteration, no symmetries
        games = 12
        args.update(workers=1, process_batch_size=12, gamesPerIteration=12, symmetricSamples=False, numMCTSSims=8, numFastSims=3,
                    arenaCompare=2, arena_batch_size=1)

    class C(GpuSelfPlayMixin, coach_mod.Coach):
        pass

    seen = []
    orig = GpuSelfPlayMixin.processSelfPlayBatches

    def spy(self, iteration):
        orig(self, iteration)
        seen.append((iteration, self.warmup, self.games_played.value, self.file_queue.qsize(), type(self.file_queue)))
    monkeypatch.setattr(GpuSelfPlayMixin, "processSelfPlayBatches", spy)

    torch.manual_seed(0)
    nnet = NNetWrapper(Game, args)
    c = C(Game, nnet, args)
    c.learn()                                        # the reference's loop, unmodified

    assert [s[:2] for s in seen] == [(1, True), (2, False)]        # warm-up iteration, then the network in the loop
    assert all(s[2] == games and s[3] > games * 5 and s[4] is ExampleQueue for s in seen)
    assert c.self_play_iter >= 1 and c.model_iter == 3
    for it in (1, 2):
        base = os.path.join(args.data, "dropin", f"iteration-{it:04d}")
        d, p, v = (torch.load(base + s) for s in ("-data.pkl", "-policy.pkl", "-value.pkl"))
        n = d.shape[0]
        assert d.shape == (n,) + obs and p.shape == (n, A) and v.shape == (n, 3) and n == seen[it - 1][3]
        assert torch.allclose(p.sum(1), torch.ones(n), atol=1e-5) and bool((v.sum(1) == 1).all())
        assert os.path.exists(os.path.join(args.checkpoint, "dropin", f"iteration-{it:04d}.pkl"))      # Coach.train ran on them
    assert np.isfinite(c.loss_pi) and np.isfinite(c.loss_v) and c.loss_pi > 0


def test_mixin_runs_workers_times_batch_games_at_once(tmp_path, monkeypatch):
    """B = workers x process_batch_size concurrent games, as Coach.generateSelfPlayAgents starts (Coach.py:294-323)."""
    coach_mod = _ref()
    from alphazero.envs.connect4.connect4 import Game
    from azb200 import coach as azcoach
    made = []
    orig = azcoach.SelfPlayEngine

    def spy(*a, **kw):
        made.append(kw["num_games"])
        return orig(*a, **kw)
    monkeypatch.setattr(azcoach, "SelfPlayEngine", spy)
    args = _args(coach_mod, tmp_path, workers=3, process_batch_size=32, gamesPerIteration=400)
    res = azcoach.run_selfplay_iteration(Game, None, args, warmup=True)
    assert made == [96] and len(res.result_turns) >= 400


def test_real_arena_play_games_with_the_import_swap_equals_the_reference_agent(tmp_path, monkeypatch):
    """Arena.play_games(use_batched_mcts=True) (Arena.pyx:208-328), run twice on fed seeds: with the reference's own
    SelfPlayAgent process and after the one-line import swap of INTEGRATION.md (alphazero.Arena.SelfPlayAgent =
    azb200.arena.ArenaAgent).  One agent with one game at a time (a slot of the engine is a batch-1 reference agent);
    the agent's stream is seeded with 777 in both runs (np.random.seed() in SelfPlayAgent.run is fed), the fast-move coin
    is fed (never fast) and the side assignment is the reference's.  Same wins per player, draws and winrates."""
    coach_mod = _ref()
    import alphazero.Arena as arena_mod
    from alphazero.GenericPlayers import MCTSPlayer
    from alphazero.NNetWrapper import NNetWrapper
    from alphazero.envs.connect4.connect4 import Game
    from azb200 import arena as azarena
    monkeypatch.chdir(tmp_path)
    args = _args(coach_mod, tmp_path, workers=1, arena_batch_size=1, numMCTSSims=12, numFastSims=3, probFastSim=0.5,
                 arenaTemp=0.25)
    torch.manual_seed(1)
    nets = [NNetWrapper(Game, args), NNetWrapper(Game, args)]
    players = [MCTSPlayer(n, Game, args) for n in nets]
    N = 6

    seed_fn = np.random.seed
    seed_fn(5)                                                              # the parent draws player_to_index (SelfPlayAgent.pyx:44-46)
    monkeypatch.setattr(np.random, "seed", lambda *a, **k: seed_fn(777))    # SelfPlayAgent.run reseeds: fed
    monkeypatch.setattr(np.random, "random_sample", lambda *a, **k: 1.0)    # the fast-move coin: fed, never fast
    ref_arena = arena_mod.Arena(players, Game, use_batched_mcts=True, args=args)
    ref_out = ref_arena.play_games(N)
    p2i = list(ref_arena._agents[0].player_to_index)

    made = []

    class SwappedAgent(azarena.ArenaAgent):
        def __init__(self, id, game_cls, *a, **kw):
            eng = azarena.arena_engine(game_cls, a[11], 1, rng="mt19937", mt_seeds=[777, 777])    # a[11] = args
            super().__init__(id, game_cls, *a, engine=eng, **kw)
            self.player_to_index = list(p2i)
            made.append(self)

        def _coin(self):
            return 1.0                                                      # fed, as the reference's
    monkeypatch.setattr(arena_mod, "SelfPlayAgent", SwappedAgent)           # from azb200.arena import ArenaAgent as SelfPlayAgent
    our_arena = arena_mod.Arena(players, Game, use_batched_mcts=True, args=args)
    our_out = our_arena.play_games(N)
    assert len(made) == 1
    assert list(our_out[0]) == list(ref_out[0]) and our_out[1] == ref_out[1], (our_out, ref_out)
    assert np.allclose(our_out[2], ref_out[2])
    assert sum(ref_out[0]) + ref_out[1] == N
    made[0].engine.close()
