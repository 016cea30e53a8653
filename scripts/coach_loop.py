#!/usr/bin/env python
"""BASELINE config 5: the Connect4 Coach loop (self-play + train + arena gating) on the engine.
  python scripts/coach_loop.py [--iters 3] [--games 8192]                      one GPU
  python -m torch.distributed.run --nproc-per-node N scripts/coach_loop.py     self-play sharded over N GPUs
Prints one JSON line per iteration (rank 0)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
import torch


class Connect4:
    """Static surface of the Connect4 plugin the loop reads (alphazero/envs/connect4/connect4.pyx:20-32)."""
    __module__ = "alphazero.envs.connect4.connect4"
    observation_size = staticmethod(lambda: (4, 6, 7))
    action_size = staticmethod(lambda: 7)
    num_players = staticmethod(lambda: 2)
    max_turns = staticmethod(lambda: 42)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--games", type=int, default=8192, help="gamesPerIteration (all ranks together)")
    ap.add_argument("--batch", type=int, default=8192, help="concurrent games per GPU")
    ap.add_argument("--sims", type=int, default=100)
    ap.add_argument("--arena", type=int, default=128)
    ap.add_argument("--graph-train", action="store_true", help="replay the training step as a CUDA graph (3.3 vs 4 ms per step)")
    ap.add_argument("--ddp-train", nargs="?", const=True, default=False, help="N > 1: every rank trains; `--ddp-train local` = per-rank window shards (no example gather), bare flag = rank 0 draws and broadcasts each batch")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from azb200.loop import GpuCoach
    coach = GpuCoach(Connect4, dict(numIters=a.iters, gamesPerIteration=a.games, process_batch_size=a.batch, numMCTSSims=a.sims,
                                    arenaCompare=a.arena, ddp_train=a.ddp_train, graph_train=a.graph_train), device=local)
    for rec in coach.learn():
        if coach.rank == 0:
            print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}))
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
