#!/usr/bin/env python
"""What the host-tensor agent protocol costs in transfers alone: per agent and simulation four DEPENDENT PCIe copies
(observations device -> pinned host -> device, answers device -> pinned host -> device), no kernels in between.
One process per GPU (torchrun) so the ranks load the host's PCIe / DRAM together, `agents` chains in flight per rank.
Prints one JSON line: per-rank GB/s each way and the simulations/s the chain alone would allow.
  python -m torch.distributed.run --nproc-per-node N scripts/pcie_chain_probe.py [--agents 4] [--games 8192]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--games", type=int, default=8192)
    ap.add_argument("--sims", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=10)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from azb200.nnet import capture_graph, upload
    Bw = a.games // a.agents
    chains = []
    for _ in range(a.agents):
        d_obs, d_in = torch.zeros(Bw, 4, 6, 7, device=dev), torch.zeros(Bw, 4, 6, 7, device=dev)
        d_ans, d_back = torch.zeros(Bw, 10, device=dev), torch.zeros(Bw, 10, device=dev)
        h_obs, h_ans = torch.zeros(Bw, 4, 6, 7).pin_memory(), torch.zeros(Bw, 10).pin_memory()
        s = torch.cuda.Stream(device=dev)

        def body(d_obs=d_obs, d_in=d_in, d_ans=d_ans, d_back=d_back, h_obs=h_obs, h_ans=h_ans):
            for _ in range(a.sims):
                h_obs.copy_(d_obs, non_blocking=True)        # generateBatch: observations -> batch_tensor
                upload(d_in, h_obs)                          # NNetWrapper.process: batch.cuda()
                h_ans.copy_(d_ans, non_blocking=True)        # policy_tensor / value_tensor .copy_(answers)
                upload(d_back, h_ans)                        # processBatch: answers -> engine
        with torch.cuda.stream(s):
            body()
            s.synchronize()
        chains.append((capture_graph(body, s), s))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(a.rounds):
        for g, s in chains:
            with torch.cuda.stream(s):
                g.replay()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dt = float(t.item())
    each_way = a.rounds * a.sims * a.games * (4 * 6 * 7 + 10) * 4 / dt / 1e9
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"n_gpus": world, "agents_per_gpu": a.agents, "games_per_gpu": a.games, "seconds": dt,
                          "pcie_gbs_each_way_per_gpu": each_way, "host_dram_gbs_all_gpus": 2 * each_way * world,
                          "chain_only_sims_per_s_per_gpu": a.rounds * a.sims * a.games / dt,
                          "chain_only_sims_per_s_all_gpus": world * a.rounds * a.sims * a.games / dt}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
