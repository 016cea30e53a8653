"""How many of the leaves a simulation round sends to the network are duplicates of another game's leaf?
python scripts/dup_probe.py [connect4|brandubh] [games] [sims] [preroll rounds] [move-rounds]
(identical observations get identical answers from the evaluator: a measure of what leaf de-duplication could save)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
from azb200 import SelfPlayEngine, default_temp_scaling, temp_table  # noqa: E402
from azb200 import nnet as aznet, nn_tc  # noqa: E402

game = sys.argv[1] if len(sys.argv) > 1 else "connect4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 100
preroll = int(sys.argv[4]) if len(sys.argv) > 4 else 48
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 3
tafl = game == "brandubh"
OBS, A = ((5, 7, 7), 588) if tafl else ((4, 6, 7), 7)
dev = torch.device("cuda")
eng = SelfPlayEngine(game=game, num_games=B, device=0, rng="philox", seed=0, cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1,
                     root_policy_temp=1.1, add_root_noise=True, add_root_temp=True, symmetric_samples=True,
                     games_per_iteration=0, max_sims_per_move=sims, max_nodes_per_game=0, sample_capacity=(600000 if tafl else 0),
                     temps=temp_table(default_temp_scaling, 1, None if tafl else 42))
torch.manual_seed(0)
model = aznet.ResNet(OBS, A, 3, **(aznet.BRANDUBH_TRAIN_NET_ARGS if tafl else aznet.DEFAULT_NET_ARGS)).to(dev).eval()
ev = nn_tc.make_evaluator(model, eng.obs, eng.policy, eng.value, rows=eng.nn_rows, count=eng.nn_count_ptr, max_batch=B)
scratch = None


def clear():
    global scratch
    n = eng.sample_count()
    if n:
        if scratch is None or scratch[0].shape[0] < n:
            m = int(n * 1.5) + 1024
            scratch = (torch.empty((m,) + OBS, device=dev), torch.empty(m, A, device=dev), torch.empty(m, 3, device=dev))
        eng.drain_samples_into(*scratch)


for i in range(preroll):
    eng.warmup_sims(8)
    eng.play_moves(False)
    if i % 4 == 3:
        clear()
clear()
for r in range(rounds):
    tot = uniq = 0
    per = []
    for s in range(sims):
        eng.select(0, B)
        cnt = int(eng._wrap(eng.nn_count_ptr(), (1,), "<i4").item())
        rows = eng.nn_rows[:cnt].long()
        o = eng.obs[rows].reshape(cnt, -1)
        u = torch.unique(o, dim=0).shape[0]
        tot += cnt; uniq += u
        if s in (0, 1, 2, 5, 10, 50, sims - 1):
            per.append((s, cnt, u))
        ev()
        eng.expand_backup(0, B)
    eng.play_moves(False)
    clear()
    print(f"{game} {B} games, move-round {r} after {preroll} pre-roll rounds: {tot} leaves to the network, {uniq} distinct "
          f"({100.0 * (1 - uniq / max(tot, 1)):.1f} % duplicates); (sim, leaves, distinct): {per}", flush=True)
