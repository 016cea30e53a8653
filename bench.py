#!/usr/bin/env python
"""bench.py -- self-play MCTS simulations/sec, 8192 Connect4 games per B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step is one move-round of the hot path for every game of the rank:
``sims`` x (select -> ResNet leaf evaluation -> expand/backup) + playMoves
(BASELINE.json configs[1]: Connect4, 8192 concurrent games, 100 sims/move,
DEFAULT_ARGS net and MCTS hyper-parameters, root noise + temperature on).
Rank r owns games [r*8192, (r+1)*8192) -- no data-path collective ("weak").

value     device-resident throughput: NN reads/writes the engine buffers in HBM.
e2e       same metric through the reference-facing SelfPlayAgent surface with
          HOST batch/policy/value tensors (Coach.processSelfPlayBatches' loop
          body), host<->device copies and sample drains inside the timed region.
e2e_coach (N = 1, secondary) the self-play phase as the drop-in Coach calls it
          (azb200.coach.run_selfplay_iteration): weights from host memory, the
          iteration's examples and results back in host memory; wall clock.
roofline  k_select (PUCT scan): algorithmic bytes 16*sumD + 12*sumC from the
          engine's exact counters / CUDA-event time of the select launches.
roofline_nn  the leaf evaluator against the measured sustained bf16 tensor
          throughput: useful flops of the ResNet x rows evaluated / CUDA-event
          time of its launches (same eager rounds as `roofline`).
cpu_baseline / --impl reference
          the reference's own Cython SelfPlayAgent processes (oracle/_ref) served
          by the reference's ResNet as Coach.processSelfPlayBatches does, on
          this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))

METRIC = "self-play MCTS simulations/sec (8192 Connect4 games)"
UNIT = "sims/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--game", default="connect4", choices=["connect4", "brandubh"],
                    help="brandubh = BASELINE config 4 (use --games 4096 --sims 200 --net brandubh_train)")
    ap.add_argument("--games", type=int, default=8192)
    ap.add_argument("--sims", type=int, default=100)
    ap.add_argument("--net", default="default", choices=["default", "connect4_train", "brandubh_train"])
    ap.add_argument("--precision", default=None, choices=["bf16x2", "fp16x2", "fp16", "bf16", "fp32", "tf32", "cudnn-bf16"],
                    help="operand precision of the leaf evaluator (accumulation is fp32): bf16x2 = hi+lo bf16 operands, within 1e-5 "
                         "of the fp32 module (default of --nn tc); fp16 / bf16 = one-pass modes of the same kernels; fp32 / tf32 / "
                         "cudnn-bf16 = PyTorch/cuDNN (--nn cudnn; default tf32, the reference's own arithmetic)")
    ap.add_argument("--cohorts", type=int, default=1)
    ap.add_argument("--split", type=int, default=0, help="games in the first cohort (0 = automatic)")
    ap.add_argument("--nn", default="tc", choices=["tc", "tc-r1", "mma", "cudnn", "fused", "fused_mma"],
                    help="leaf evaluator: tc = hand-written tcgen05/TMEM kernels (csrc/azb_resnet_g.cu, every shipped geometry: 32 / 64 / 128 "
                         "channels); tc-r1 / mma = the round-1 bf16-only kernels (6x7, 32 channels); cudnn = PyTorch/cuDNN CUDA graph")
    ap.add_argument("--lanes", type=int, default=0, help="threads per game (0 = library default)")
    ap.add_argument("--nchw", action="store_true", help="keep the ResNet in NCHW (default: channels_last)")
    ap.add_argument("--preroll", type=int, default=48, help="cheap tree-only rounds that de-synchronise the games")
    ap.add_argument("--tree-only", action="store_true", help="warmup mode (constant NN outputs), one fused kernel per round")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--sustain-seconds", type=float, default=5.0, help="extra sustained leg of the device-resident step (0 = off)")
    ap.add_argument("--e2e-agents", type=int, default=4, help="reference `workers`: agents sharing the GPU in the e2e leg")
    ap.add_argument("--e2e-mode", default="inline", choices=["inline", "threads"],
                    help="inline: one host thread runs the Coach loop over the agents (generateBatch -> process -> "
                         "processBatch, every call asynchronous and stream-ordered); threads: agents are threads, ready queue")
    ap.add_argument("--e2e-eager", action="store_true", help="inline e2e leg without the agents' captured step graphs")
    ap.add_argument("--e2e-step-graphs", action="store_true",
                    help="inline e2e leg with three step graphs per agent and simulation (round 1) instead of one graph per agent "
                         "and move-round with the server body inside")
    ap.add_argument("--e2e-item-queue", action="store_true",
                    help="e2e leg: the agents put one queue item per example (reference file_queue protocol) instead of blocks")
    ap.add_argument("--e2e-sync", action="store_true",
                    help="e2e leg with host-synchronised tensors (default: stream-ordered, see SelfPlayAgent)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true")
    ap.add_argument("--leaf-dedup", action="store_true",
                    help="headline with leaf de-duplication on (default: every non-terminal leaf is evaluated; the "
                         "de-duplicated step is reported as the extra key leaf_dedup)")
    ap.add_argument("--no-select-events", action="store_true")
    ap.add_argument("--roofline-steps", type=int, default=3, help="eager move-rounds timed per select launch for the roofline")
    ap.add_argument("--no-round-graph", action="store_true", help="launch every kernel of a move-round from the host")
    ap.add_argument("--eval-terminal", action="store_true",
                    help="evaluate the network for terminal leaves too, as the reference does (it discards those answers)")
    return ap.parse_args()


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed region (nvidia-smi itself takes about a second to come up, so the
        sampler is started long before; only samples taken after this mark are reported)."""
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        if sum(1 for ts, _ in self.lines if ts >= t_mark) < 3:
            # short timed region: count the warm-up steps too (same load, run immediately before)
            t_mark = getattr(self, "t_warm", t_mark)
        for ts, ln in self.lines:
            if ts < t_mark:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# reference arm: the reference's own SelfPlayAgent processes + NN server loop
# --------------------------------------------------------------------------
def run_reference(games, sims, net, steps, warmup, seconds=None, use_cuda=True):
    """Times the compiled reference (oracle/_ref).  Workers are the reference's
    mp.Process agents; this process is the NN server exactly as
    Coach.processSelfPlayBatches (Coach.py:326-361).  One step = one served
    batch per worker x 8.  Returns a dict or {'unavailable': why}."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref
        if not build_ref.built():
            return {"unavailable": "oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)"}
    finally:
        sys.path.pop(0)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import numpy as np  # noqa: F401
    import torch
    import torch.multiprocessing as mp
    from queue import Empty
    from alphazero.SelfPlayAgent import SelfPlayAgent
    from alphazero.envs.connect4.connect4 import Game
    from alphazero.NNetArchitecture import ResNet
    from alphazero.utils import dotdict, default_temp_scaling
    from azb200 import nnet as aznet

    cores = os.cpu_count() or 2
    workers = max(1, min(cores - 1, games))
    pbs = -(-games // workers)
    netargs = aznet.DEFAULT_NET_ARGS if net == "default" else aznet.CONNECT4_TRAIN_NET_ARGS
    args = dotdict(dict(
        startTemp=1, temp_scaling_fn=default_temp_scaling, root_noise_frac=0.1, root_policy_temp=1.1,
        min_discount=1, fpu_reduction=0.2, cpuct=1.25, _num_players=2, add_root_noise=True, add_root_temp=True,
        symmetricSamples=True, mctsResetThreshold=None, gamesPerIteration=1 << 40, numMCTSSims=sims,
        numFastSims=sims, numWarmupSims=sims, probFastSim=0.0, arenaTemp=0.25, process_batch_size=pbs,
        workers=workers, **netargs))
    torch.manual_seed(0)
    model = ResNet(Game, args)
    stop, pause = mp.Event(), mp.Event()
    ready, fileq, resq = mp.Queue(), mp.Queue(), mp.Queue()
    completed, played = mp.Value('i', 0), mp.Value('i', 0)
    ins, pols, vals, evs, agents = [], [], [], [], []
    for i in range(workers):
        ins.append(torch.zeros([pbs, *Game.observation_size()]).share_memory_())
        pols.append(torch.zeros([pbs, Game.action_size()]).share_memory_())
        vals.append(torch.zeros([pbs, 3]).share_memory_())
        evs.append(mp.Event())
        ag = SelfPlayAgent(i, Game, ready, evs[i], ins[i], pols[i], vals[i], fileq, resq, completed, played,
                           stop, pause, args)
        ag.daemon = True
        agents.append(ag)
    for ag in agents:      # fork before this process touches CUDA
        ag.start()
    cuda = use_cuda and torch.cuda.is_available()
    if cuda:
        model.cuda()
    model.eval()

    def process(batch):   # NNetWrapper.process (NNetWrapper.py:225-232)
        batch = batch.type(torch.FloatTensor)
        if cuda:
            batch = batch.cuda()
        with torch.no_grad():
            pi, v = model(batch)
            return torch.exp(pi), torch.exp(v)

    served = 0

    def serve(n_batches, deadline=None):
        nonlocal served
        done = 0
        while done < n_batches and (deadline is None or time.time() < deadline):
            try:
                i = ready.get(timeout=1)
            except Empty:
                continue
            p, v = process(ins[i])
            pols[i].copy_(p)
            vals[i].copy_(v)
            evs[i].set()
            done += 1
            served += 1
            # keep the queues from filling the pipes
            for q in (fileq, resq):
                try:
                    while True:
                        q.get_nowait()
                except Empty:
                    pass
        return done

    per_step = workers * 8
    serve(per_step * max(warmup, 1))
    t0 = time.time()
    if seconds is not None:
        n = serve(1 << 60, deadline=t0 + seconds)
        nsteps = max(1, n // per_step)
    else:
        n = serve(per_step * steps)
        nsteps = steps
    if cuda:
        torch.cuda.synchronize()
    dt = time.time() - t0
    stop.set()
    for e in evs:
        e.set()
    t_end = time.time() + 20
    for ag in agents:
        for q in (ready, fileq, resq):
            try:
                while True:
                    q.get_nowait()
            except Empty:
                pass
        ag.join(timeout=max(0.1, t_end - time.time()))
        if ag.is_alive():
            ag.terminate()
    total_sims = n * pbs
    return {"value": total_sims / dt, "seconds": dt, "sims": total_sims, "cores": workers + 1, "workers": workers,
            "process_batch_size": pbs, "steps": nsteps, "nn_device": "cuda" if cuda else "cpu",
            "sample": f"{n} NN-served batches of {pbs} games ({total_sims} simulations, {dt:.1f} s) by {workers} "
                      f"reference SelfPlayAgent processes + 1 NN-server process; ResNet on {'cuda' if cuda else 'cpu'}"}


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_reference(a.games, a.sims, a.net, a.steps, a.warmup)
    if "unavailable" in r:
        print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}))
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": r["steps"],
        "warmup": a.warmup, "ms_per_step": 1000.0 * r["seconds"] / r["steps"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"connect4 {a.games} games x {a.sims} sims/move, net={a.net}, reference Cython "
                               f"SelfPlayAgent x{r['workers']} workers (batch {r['process_batch_size']})"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
def main():
    a = parse()
    wd = float(os.environ.get("AZB_BENCH_WATCHDOG", "0"))
    if wd > 0:      # debugging aid: dump every thread's stack and exit if the run exceeds this many seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    if a.impl == "reference":
        reference_main(a)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.game == "connect4":
        # before this process initialises CUDA: the reference agents are forked
        try:
            r = run_reference(a.games, a.sims, a.net, 0, 1, seconds=a.cpu_seconds)
            if "unavailable" not in r:
                cpu_base = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                            "sample": r["sample"]}
            else:
                cpu_base = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": r["unavailable"]}
        except Exception as ex:   # the baseline must never take the bench down
            cpu_base = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex!r}"}

    import numpy as np
    import torch
    import torch.distributed as dist
    from azb200 import SelfPlayEngine, default_temp_scaling, temp_table
    from azb200 import nnet as aznet
    from azb200.selfplay import DeviceSelfPlay, SelfPlayAgent

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    B, sims = a.games, a.sims
    tafl = a.game == "brandubh"
    OBS, A = ((5, 7, 7), 588) if tafl else ((4, 6, 7), 7)
    eng = SelfPlayEngine(
        game=a.game, num_games=B, device=local, rng="philox", seed=0, game_id_base=rank * B,
        cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1, add_root_noise=True,
        add_root_temp=True, symmetric_samples=True, games_per_iteration=0, max_sims_per_move=sims,
        max_nodes_per_game=0, sample_capacity=(600000 if tafl else 0),
        temps=temp_table(default_temp_scaling, 1, None if tafl else 42), lanes_per_game=a.lanes)
    torch.manual_seed(0)
    netargs = {"default": aznet.DEFAULT_NET_ARGS, "connect4_train": aznet.CONNECT4_TRAIN_NET_ARGS,
               "brandubh_train": aznet.BRANDUBH_TRAIN_NET_ARGS}[a.net]
    model = aznet.ResNet(OBS, A, 3, **netargs).to(dev).eval()
    from azb200 import nn_tc
    a.nn = {"fused": "tc", "fused_mma": "mma"}.get(a.nn, a.nn)
    if a.nn == "tc" and not nn_tc.supported(model):
        a.nn = "cudnn"                      # a geometry the hand-written kernels do not cover (boards > 7x7)
    if tafl:
        a.no_e2e = True                     # the host-tensor agent legs below are written for Connect4 shapes
    if a.nn == "cudnn":
        a.precision = a.precision if a.precision in ("fp32", "tf32", "cudnn-bf16") else "tf32"
    elif a.nn == "tc":
        a.precision = a.precision if a.precision in nn_tc.PRECISIONS else nn_tc.default_precision(model)
    else:
        a.precision = "bf16"
    drv = DeviceSelfPlay(eng, model, cohorts=a.cohorts, precision=a.precision, channels_last=not a.nchw,
                         fused={"tc": None, "tc-r1": "tc-r1", "mma": "mma", "cudnn": False}[a.nn],
                         round_graph=False if a.no_round_graph else None, split=a.split or None,
                         skip_terminal=not a.eval_terminal, dedup=a.leaf_dedup)
    compact = a.nn in ("tc", "tc-r1") and not a.eval_terminal and a.cohorts == 1
    nn_kernel = {"tc": ("k_trunk_wide" if netargs["num_channels"] == 128 else "k_trunk_tc") + " + k_head_tc (tcgen05/TMEM, csrc/azb_resnet_g.cu)", "tc-r1": "k_resnet_tc (tcgen05/TMEM, round 1)",
                 "mma": "k_resnet_fused (mma.sync)"}.get(a.nn, "cuDNN " + a.precision)
    # what the arithmetic is: operand type of the convolutions / head GEMM + accumulator type
    dtype = {"bf16x2": "bf16x2-split operands (16 significant bits) + f32 accumulate",
             "fp16x2": "f16x2-split operands (22 significant bits) + f32 accumulate", "fp16": "f16 operands + f32 accumulate",
             "bf16": "bf16 operands + f32 accumulate", "cudnn-bf16": "bf16 (autocast)", "tf32": "tf32", "fp32": "f32"}[a.precision]

    sel_events, nn_events = [], []

    def step():
        if a.tree_only:
            drv.run_round_warmup(sims)
        else:
            drv.run_round(sims)
        # keep the sample ring from filling: discard on the device (no host copy)
        clear_samples()

    # Examples are drained on the device into buffers allocated ONCE (a fresh torch.empty per step that is never
    # freed means a cudaMalloc -- a device-wide synchronisation -- inside the timed region): a scratch triple when the
    # examples are discarded (N = 1), a keep buffer for the NCCL gather at the end (N > 1).
    keep_cap = 1_200_000 if world > 1 else 0
    keep = [torch.empty((keep_cap,) + OBS, device=dev), torch.empty(keep_cap, A, device=dev), torch.empty(keep_cap, 3, device=dev)]
    scratch = [None]
    kept = [0]

    def clear_samples():
        n = eng.sample_count()
        if n == 0:
            return
        k = kept[0]
        if k + n <= keep_cap:
            eng.drain_samples_into(keep[0][k:k + n], keep[1][k:k + n], keep[2][k:k + n])
            kept[0] = k + n
            return
        if scratch[0] is None or scratch[0][0].shape[0] < n:
            m = int(n * 1.5) + 1024
            scratch[0] = (torch.empty((m,) + OBS, device=dev), torch.empty(m, A, device=dev), torch.empty(m, 3, device=dev))
        eng.drain_samples_into(*scratch[0])

    # cheap tree-only pre-roll so the games are spread over all phases (steady state)
    for _ in range(a.preroll):
        eng.warmup_sims(8)
        eng.play_moves(False)
        if _ % 4 == 3:
            clear_samples()
    clear_samples()
    kept[0] = 0
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks.t_warm = time.time()
    for _ in range(max(a.warmup, 3)):
        step()
    clear_samples()
    kept[0] = 0
    eng.check_errors()
    st0 = eng.stats()
    launches0 = drv.launches

    barrier()
    clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    st1 = eng.stats()
    eng.check_errors()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    dsims = st1["sims"] - st0["sims"]
    tot = torch.tensor([dsims], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_sims = float(tot.item())
    value = total_sims / (ms_max / 1000.0)

    # roofline of the PUCT selection kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    dD, dC = st1["sum_depth"] - st0["sum_depth"], st1["sum_children"] - st0["sum_children"]
    alg_bytes = 16.0 * dD + 12.0 * dC
    roof = None
    # Roofline pass: the timed steps replay one CUDA graph per move-round, so the per-launch time of k_select is taken
    # right after them from eagerly launched rounds of the same workload, with CUDA events around every select launch
    # on the tree stream (the kernel and its inputs are the same; only the launch path differs).
    if not a.no_select_events and not a.tree_only:
        orig_select, was_graph = eng.select, drv.round_graph
        drv.round_graph = False

        def timed_select(first=0, count=0, stream=None):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            orig_select(first, count, stream=stream)
            e1.record(stream)
            sel_events.append((e0, e1))
        eng.select = timed_select
        orig_evals = list(drv.evals)

        def timed_eval(ev):
            def call(stream=None):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ev(stream=stream)
                e1.record(stream)
                nn_events.append((e0, e1))
            return call
        drv.evals = [timed_eval(ev) for ev in orig_evals]
        r0 = eng.stats()
        for _ in range(a.roofline_steps):
            step()
        torch.cuda.synchronize()
        r1 = eng.stats()
        eng.select, drv.round_graph, drv.evals = orig_select, was_graph, orig_evals
        alg_bytes = 16.0 * (r1["sum_depth"] - r0["sum_depth"]) + 12.0 * (r1["sum_children"] - r0["sum_children"])
        roof_sims = r1["sims"] - r0["sims"]
    def launch_times(events):
        """CUDA-event time per launch.  In these eagerly launched rounds the start event can execute before the host
        has issued the kernel behind it (a Python / scheduler pause between the two calls then counts as kernel time),
        so launches longer than 3x the median are set aside; their number and the untrimmed mean are reported."""
        ts = [e0.elapsed_time(e1) for e0, e1 in events]
        med = statistics.median(ts)
        kept_ts = [t for t in ts if t <= 3.0 * med]
        q = sorted(ts)
        launch_times.last_percentiles = [1000.0 * q[int(f * (len(q) - 1))] for f in (0.1, 0.5, 0.9)]
        return sum(kept_ts), len(kept_ts), len(ts) - len(kept_ts), 1000.0 * sum(ts) / len(ts)

    if sel_events:
        sel_ms, nl, sel_dropped, sel_mean_all = launch_times(sel_events)
        sel_pct = launch_times.last_percentiles
        alg_bytes *= nl / float(len(sel_events))
        ach = alg_bytes / (sel_ms / 1000.0) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": None, "kernel": f"k_select<{'Brandubh' if tafl else 'Connect4'}>", "launches": nl, "avg_launch_us": 1000.0 * sel_ms / nl,
                "alg_bytes_per_launch": alg_bytes / nl, "bytes_per_sim": alg_bytes * (len(sel_events) / float(nl)) / max(roof_sims, 1), "peak_source": peak_src,
                "launches_set_aside": sel_dropped, "mean_launch_us_untrimmed": sel_mean_all,
                "launch_us_p10_p50_p90": sel_pct,
                "note": f"CUDA-event time of each select launch on the tree stream over {a.roofline_steps} eagerly launched "
                        "move-rounds right after the timed (graph-replayed) steps"}
    elif a.tree_only:
        ach = alg_bytes / (ms / 1000.0) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": None, "kernel": "k_warmup_sims<Connect4> (select+expand+backup fused)", "launches": a.steps,
                "avg_launch_us": 1000.0 * ms / a.steps, "alg_bytes_per_launch": alg_bytes / a.steps,
                "bytes_per_sim": alg_bytes / max(dsims, 1), "peak_source": peak_src}
    # secondary roofline: the leaf evaluator (the kernel that takes most of the step) against the measured dense bf16
    # tensor throughput; flops = conv / linear multiply-adds of the ResNet x the rows it evaluated (non-terminal leaves
    # in compact mode, every leaf otherwise), time = CUDA events around its launches in the same eager rounds
    roof_nn = None
    if nn_events:
        nn_ms, nn_n, nn_dropped, nn_mean_all = launch_times(nn_events)
        nn_pct = launch_times.last_percentiles
        rows = (roof_sims - ((r1["terminal_leaves"] - r0["terminal_leaves"]) if compact else 0)) * (nn_n / float(len(nn_events)))
        fl = model_flops(model, OBS)
        tf = fl * rows / (nn_ms / 1000.0) / 1e12
        sustained = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1388.0)))
        roof_nn = {"bound": "tensor", "achieved": tf, "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained, "traffic": None,
                   "kernel": nn_kernel, "precision": a.precision,
                   "launches": nn_n, "avg_launch_us": 1000.0 * nn_ms / nn_n, "flops_per_eval": fl,
                   "rows_per_launch": rows / nn_n, "evals_per_s": rows / (nn_ms / 1000.0),
                   "launches_set_aside": nn_dropped, "mean_launch_us_untrimmed": nn_mean_all, "launch_us_p10_p50_p90": nn_pct,
                   "peak_source": ("measured sustained bf16" if "bf16_tflops_sustained" in peaks else "fallback"),
                   "note": "useful flops of the network (2 x multiply-adds of its convolutions and linear layers), not issued MMA flops"}
        if a.nn == "tc":
            # what the tensor pipe actually executes: M = 128 x N x K = 16 MMAs over 8x8 board frames (64 rows per board for
            # H x W live positions), three per K step at a split precision; trunk only (the head GEMM is < 2 %)
            ch, dep = netargs["num_channels"], netargs["depth"]
            passes = 3 if a.precision in nn_tc.SPLIT else 1
            if ch == 128:
                mmas, n_mma = (3 * 2 + 2 * dep * 36 * 2) * passes, 128
            else:
                mmas, n_mma = (2 + 2 * dep * 3 * (ch // 16)) * passes, 3 * ch
            issued = (rows / nn_n / 2.0) * mmas * 2.0 * 128 * n_mma * 16 / (nn_ms / nn_n / 1000.0) / 1e12
            roof_nn["issued_mma"] = {"tflops": issued, "frac_of_sustained_peak": issued / sustained, "mma_per_tile": mmas, "mma_n": n_mma,
                                     "note": "issued tensor-core flops of the trunk kernel (frames of 64 rows per board, "
                                             f"{passes} MMA(s) per K step) over the same time: the split precision and the frame padding "
                                             "are what separate it from the useful-flop figure"}
    traffic_file = os.path.join(ROOT, "profiles", "select_traffic.json")
    if roof is not None and os.path.exists(traffic_file):
        try:
            roof["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
            roof["traffic_note"] = (
                "ncu dram read+write per k_select launch with caches flushed before the launch (profiles/r2_ncu_select_summary.csv): "
                "6.45 MB read + 0.2 MB written.  Cold, every L2 read sector misses (lts__t_sectors_srcunit_tex_op_read = 200 k sectors = "
                "6.4 MB): sibling blocks 3.3 MB (7 x 16 B hot records = 112 B, of which 12 B per child are the scan's N/Q/P, spread over "
                "4-5 32-byte sectors per level), the chosen child's 8-byte cold record per level 0.7 MB (one sector each), the speculative "
                "next-block prefetch ~1 MB, slot header 0.5 MB, RNG counter and statistics rows 0.5 MB -- sector granularity around 8-16 B "
                "records, not re-reads.  Writes are 13.8 MB of L2 write sectors (leaf observations 5.5 MB, new child records, path, leaf "
                "record, header), almost none of which reach DRAM inside the launch.  In the loop (same counters with --cache-control none, "
                "profiles/r2_select_insitu_cache.csv) the launch reads 8.9 MB from DRAM and evicts 9.0 MB: the evaluator's traffic between two "
                "selects leaves none of the ~300 MB of live tree records in the 126 MB L2, every tree access is a DRAM access (17 % of the "
                "HBM peak in real bytes).  `achieved` counts only the PUCT scan (16 B header + 12 B per scanned child, SURVEY 8d)")
        except Exception:
            pass

    def timed_rounds(driver, steps, warm=2):
        for _ in range(warm):
            driver.run_round(sims); clear_samples()
        s0 = eng.stats()["sims"]
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(steps):
            driver.run_round(sims); clear_samples()
        f1.record()
        barrier()
        tt = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        nn_ = torch.tensor([eng.stats()["sims"] - s0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(nn_, op=dist.ReduceOp.SUM)
        return float(nn_.item()) / (float(tt.item()) / 1000.0), float(tt.item())

    # measured error of the evaluator the timed steps used, on leaf observations of this very run, against the PyTorch
    # module in strict fp32 (NNetWrapper.process, the oracle for this floating-point kernel)
    nn_error = None
    if not a.tree_only:
        try:
            n_chk = min(B, 2048)
            o = eng.obs[:n_chk].clone()
            pol_c, val_c = torch.zeros(n_chk, A, device=dev), torch.zeros(n_chk, 3, device=dev)
            ev_c = nn_tc.make_evaluator(model, o, pol_c, val_c, precision=a.precision,
                                        kernel=a.nn if a.nn in ("tc-r1", "mma") else None, channels_last=not a.nchw)
            ev_c()
            old_flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            with torch.no_grad():
                lp, lv = model(o)
            torch.backends.cudnn.allow_tf32 = True
            with torch.no_grad():
                lp_t, lv_t = model(o)
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_flags
            torch.cuda.synchronize()
            err = lambda p_, v_: max((p_ - lp.exp()).abs().max().item(), (v_ - lv.exp()).abs().max().item())
            nn_error = {"boards": n_chk, "max_abs_error_vs_f32_module": err(pol_c, val_c),
                        "cudnn_tf32_max_abs_error_vs_f32_module": err(lp_t.exp(), lv_t.exp()),
                        "note": "probabilities (policy and value) on leaf observations of this run; the second figure is what the "
                                "reference's own default arithmetic (PyTorch cuDNN TF32 convolutions) does on the same boards"}
            del ev_c
        except Exception as ex:
            nn_error = {"error": repr(ex)}

    # sustained leg: the same device-resident step for >= --sustain-seconds with its own clock record
    sustained_leg = None
    if a.sustain_seconds > 0 and not a.tree_only:
        n_sus = max(a.steps, int(a.sustain_seconds * 1000.0 / max(ms_max / a.steps, 1e-3)) + 1)
        clocks2 = ClockSampler(local)
        if rank == 0:
            clocks2.start(); clocks2.t_warm = time.time(); clocks2.mark()
        v_sus, ms_sus = timed_rounds(drv, n_sus, warm=0)
        sustained_leg = {"value": v_sus, "unit": UNIT, "steps": n_sus, "seconds": ms_sus / 1000.0,
                         "clocks": clocks2.stop() if rank == 0 else None}

    # the same step with leaf de-duplication (DeviceSelfPlay's default; the headline above evaluates every non-terminal leaf
    # unless --leaf-dedup): games whose leaves have the same observation share one evaluation, results bit-identical
    dedup_leg = None
    if a.nn == "tc" and compact and not a.tree_only and not a.no_alt and not a.leaf_dedup:
        d4 = DeviceSelfPlay(eng, model, cohorts=a.cohorts, precision=a.precision, skip_terminal=True, dedup=True)
        for _ in range(2):
            d4.run_round(sims); clear_samples()
        sd0, dd0 = eng.stats(), eng.duplicate_leaves()
        v4, ms4 = timed_rounds(d4, 20, warm=0)
        sd1, dd1 = eng.stats(), eng.duplicate_leaves()
        need = (sd1["sims"] - sd0["sims"]) - (sd1["terminal_leaves"] - sd0["terminal_leaves"])
        dedup_leg = {"value": v4, "unit": UNIT, "steps": 20, "ms_per_step": ms4 / 20.0,
                     "duplicate_fraction_of_non_terminal_leaves": (dd1 - dd0) / max(need, 1),
                     "note": "same step, same games, bit-identical results: a leaf whose observation equals that of another "
                             "game's leaf in the same simulation round is not evaluated again (azb_set_leaf_dedup); rank 0's "
                             "duplicate fraction"}
        del d4
        eng.set_leaf_dedup(False)

    # secondary legs, same step: the PyTorch/cuDNN TF32 evaluator (the reference's own default arithmetic) and the other
    # operand precisions of the hand-written kernels
    alt, alt_prec = None, None
    if a.nn != "cudnn" and not a.tree_only and not a.no_alt:
        if a.nn == "tc":
            alt_prec = {}
            for prec in ("bf16x2", "fp16x2", "fp16", "bf16"):
                if prec == a.precision:
                    continue
                d3 = DeviceSelfPlay(eng, model, cohorts=a.cohorts, precision=prec, skip_terminal=not a.eval_terminal, dedup=a.leaf_dedup)
                v3, _ = timed_rounds(d3, 5)
                alt_prec[prec] = {"value": v3, "unit": UNIT, "steps": 5}
                del d3
            try:
                d3 = DeviceSelfPlay(eng, model, cohorts=a.cohorts, precision="bf16", fused="tc-r1", skip_terminal=not a.eval_terminal,
                                    dedup=a.leaf_dedup)
                v3, _ = timed_rounds(d3, 5)
                alt_prec["bf16 (round-1 kernel k_resnet_tc)"] = {"value": v3, "unit": UNIT, "steps": 5}
                del d3
            except (NotImplementedError, RuntimeError):
                pass
        model.to(memory_format=torch.channels_last)
        drv2 = DeviceSelfPlay(eng, model, cohorts=a.cohorts, precision="tf32", channels_last=True, fused=False)
        v2, _ = timed_rounds(drv2, 3)
        alt = {"nn": "cudnn (CUDA graph, channels_last)", "dtype": "tf32", "value": v2, "unit": UNIT, "steps": 3}

    # e2e: reference-facing SelfPlayAgent surface with host tensors
    e2e = None
    if not a.no_e2e and not a.tree_only:
        e2e = run_e2e(a, eng, model, dev, world)

    # e2e_coach: the call a user of the drop-in Coach makes (GpuSelfPlayMixin.processSelfPlayBatches)
    e2e_coach, e2e_coach_dedup = None, None

    def coach_leg(dedup):
        # every N: the local part may fail without touching a collective; the reduction below always runs on every rank
        try:
            r = run_e2e_coach(a, model, dev, local, rank, world, dedup)
        except Exception as ex:          # a secondary number must never take the bench down
            r = {"value": None, "error": repr(ex), "seconds": float("nan"), "sims": 0.0}
        if world > 1:
            tq = torch.tensor([r["seconds"]], device=dev, dtype=torch.float64)
            nq = torch.tensor([r["sims"]], device=dev, dtype=torch.float64)
            dist.all_reduce(tq, op=dist.ReduceOp.MAX); dist.all_reduce(nq, op=dist.ReduceOp.SUM)
            if "error" not in r:
                r.update(value=float(nq.item()) / float(tq.item()), seconds=float(tq.item()), sims=float(nq.item()), ranks=world)
        return r

    if not a.no_e2e and not a.tree_only and a.game == "connect4" and a.net == "default":
        # e2e_coach keeps every non-terminal leaf evaluated (as `value`); e2e_coach_dedup is the same call with the Coach
        # path's default, leaf de-duplication -- an iteration starts all its games from the empty board together
        e2e_coach = coach_leg(a.leaf_dedup)
        if not a.leaf_dedup and not a.no_alt:
            e2e_coach_dedup = coach_leg(True)

    if world > 1:
        clear_samples()
        gather_ms, gathered = gather_examples([t[:kept[0]] for t in keep], dev, rank, world)
    else:
        gather_ms, gathered = None, None

    if rank == 0:
        line = {
            "metric": METRIC if (a.game == "connect4" and B == 8192) else f"self-play MCTS simulations/sec ({B} {a.game} games)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"{a.game} {B} games/GPU x {sims} sims/move, DEFAULT_ARGS MCTS (cpuct 1.25, fpu 0.2, "
                                   f"root noise 0.1 + temp 1.1), net={a.net} ResNet random-init, "
                                   f"{'tree-only warmup mode' if a.tree_only else 'NN in the loop'}",
                       "games_per_gpu": B, "sims_per_move": sims, "net": a.net, "nn": a.nn, "nn_kernel": nn_kernel, "nn_precision": a.precision,
                       "cohorts": a.cohorts, "round_graph": bool(drv.round_graph),
                       "nn_rows": ("distinct non-terminal leaves (leaf de-duplication)" if (compact and a.leaf_dedup) else
                                   "non-terminal leaves only" if compact else "every leaf"),
                       "nn_tile_plan": ("whole waves of one-round CTAs" if os.environ.get("AZB_NNG_PERSIST", "1") == "0"
                                        else "one wave of persistent CTAs, tiles in rounds") if a.nn == "tc" else None, "channels_last": not a.nchw, "lanes_per_game": a.lanes or 8, "rng": "philox", "parallelism": f"games x{world} (no data-path collective)",
                       "l2": "node pool %.1f GB per GPU > 126 MB L2; no flush" % (st1["pool_bytes"] / 1e9),
                       "preroll_rounds": a.preroll},
            "clocks": clk, "gpu_launches": drv.launches - launches0,
            "roofline": roof, "roofline_nn": roof_nn, "cpu_baseline": cpu_base, "e2e": e2e, "e2e_coach": e2e_coach, "e2e_coach_dedup": e2e_coach_dedup,
            "leaf_dedup": dedup_leg, "alt_nn": alt,
            "alt_precisions": alt_prec, "nn_error": nn_error, "sustained": sustained_leg,
            "tree_stats": {"sims": dsims, "mean_depth": dD / max(dsims, 1), "mean_children_scanned": dC / max(dsims, 1),
                           "games_finished": st1["results"] - st0["results"], "peak_nodes_per_game": st1["peak_nodes"],
                           "terminal_leaf_fraction": (st1["terminal_leaves"] - st0["terminal_leaves"]) / max(dsims, 1)},
        }
        if gather_ms is not None:
            line["example_gather"] = {"ms": gather_ms, "samples": gathered}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def model_flops(model, obs_shape):
    """2 x multiply-adds of every Conv2d / Linear in one evaluation of `model` (hooks over a one-board forward)."""
    import torch
    total = [0]

    def hook(m, inp, out):
        if isinstance(m, torch.nn.Conv2d):
            total[0] += 2 * out.numel() * (m.in_channels // m.groups) * m.kernel_size[0] * m.kernel_size[1]
        elif isinstance(m, torch.nn.Linear):
            total[0] += 2 * out.numel() * m.in_features
    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear))]
    try:
        p = next(model.parameters())
        with torch.no_grad():
            model(torch.zeros((1,) + tuple(obs_shape), device=p.device, dtype=p.dtype))
    finally:
        for h in hs:
            h.remove()
    return float(total[0])


def run_e2e_coach(a, model, dev, local, rank, world, dedup=False):
    """Secondary end-to-end number: the self-play phase as a user of the drop-in Coach calls it
    (azb200.coach.GpuSelfPlayMixin.processSelfPlayBatches -> run_selfplay_iteration): per iteration the network's
    weights come from HOST memory (pinned state_dict -> device), the engine plays gamesPerIteration games, and the
    training examples + game results are returned in HOST memory (what saveIterationSamples consumes).  Wall clock
    around the whole call, max over ranks."""
    import time
    import torch
    import torch.distributed as dist
    from azb200 import SelfPlayEngine
    from azb200 import nnet as aznet
    from azb200.coach import run_selfplay_iteration
    from azb200.selfplay import engine_kwargs_from_args

    class Connect4:
        __module__ = "alphazero.envs.connect4.connect4"
        observation_size = staticmethod(lambda: (4, 6, 7))
        action_size = staticmethod(lambda: 7)
        num_players = staticmethod(lambda: 2)
        max_turns = staticmethod(lambda: 42)

    B, sims = a.games, a.sims
    args = dict(process_batch_size=B, gamesPerIteration=2 * B, numMCTSSims=sims, numFastSims=sims, probFastSim=0.0,
                cpuct=1.25, fpu_reduction=0.2, root_noise_frac=0.1, root_policy_temp=1.1, add_root_noise=True,
                add_root_temp=True, symmetricSamples=True, leaf_dedup=bool(dedup))
    host_weights = {k: v.detach().cpu().pin_memory() for k, v in model.state_dict().items()}
    net = aznet.ResNet((4, 6, 7), 7, 3, **aznet.DEFAULT_NET_ARGS).to(dev).eval()
    eng = SelfPlayEngine(**engine_kwargs_from_args(Connect4, args, B, device=local, rng="philox", seed=1, game_id_base=rank * B))
    run_selfplay_iteration(Connect4, net, args, device=local, seed=1, engine=eng)   # warm-up: a previous iteration of the same size
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sd = net.state_dict()
    for k, v in host_weights.items():                                    # H2D: this iteration's network
        sd[k].copy_(v, non_blocking=True)
    res = run_selfplay_iteration(Connect4, net, args, device=local, seed=2, engine=eng)      # D2H: examples, results
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.close()
    h2d = sum(v.numel() * v.element_size() for v in host_weights.values())
    d2h = sum(x.numel() * x.element_size() for x in (res.data, res.policy, res.value)) + res.result_turns.nbytes + res.result_winstates.nbytes
    return {"value": float(res.sims) / dt, "unit": UNIT, "seconds": dt, "sims": float(res.sims), "games": int(len(res.result_turns)),
            "examples": int(res.data.shape[0]), "h2d_bytes": int(h2d), "d2h_bytes": int(d2h), "leaf_dedup": bool(dedup),
            "api": "azb200.coach.run_selfplay_iteration (the body of GpuSelfPlayMixin.processSelfPlayBatches): network weights from "
                   "pinned host memory, gamesPerIteration = 2 x games, examples and results returned in host memory; wall clock"}


def run_e2e(a, eng, model, dev, world):
    """Coach.processSelfPlayBatches (Coach.py:326-361) as the reference runs it: `workers`
    reference-shaped SelfPlayAgent objects (here threads, each owning a CUDA engine with its share
    of the games) exchange observation / policy / value batches with this NN-server loop through
    pinned HOST tensors, a ready queue and per-agent events."""
    import queue as pyqueue
    import torch
    import torch.distributed as dist
    from azb200 import SelfPlayEngine, default_temp_scaling, temp_table
    from azb200.nnet import NNetWrapper
    from azb200.selfplay import SelfPlayAgent

    class _Val:
        def __init__(self): self.value = 0; self._l = threading.Lock()
        def get_lock(self): return self._l

    class _Sink:
        def put(self, x): pass

    class _Game:
        __module__ = "alphazero.envs.connect4.connect4"
    W = max(1, a.e2e_agents)
    Bw = eng.B // W
    rank = int(os.environ.get("RANK", "0"))
    args = type("A", (dict,), {"__getattr__": dict.__getitem__})(
        gamesPerIteration=1 << 40, probFastSim=0.0, numMCTSSims=a.sims, numFastSims=a.sims, numWarmupSims=a.sims)
    ready, stop, pause = pyqueue.Queue(), threading.Event(), threading.Event()
    completed, played = _Val(), _Val()
    bts, pts, vts, evs, agents = [], [], [], [], []
    # Coach.file_queue: azb200.coach.ExampleQueue (what GpuSelfPlayMixin installs) keeps every example, block-wise;
    # --e2e-item-queue: an mp.Queue-like sink that takes one (obs, pi, z) item per example, as the reference's does
    from azb200.coach import ExampleQueue
    file_queue = _Sink() if a.e2e_item_queue else ExampleQueue()
    for i in range(W):
        e = SelfPlayEngine(game="connect4", num_games=Bw, device=dev.index, rng="philox", seed=1,
                           game_id_base=(rank * W + i) * Bw, add_root_noise=True, add_root_temp=True,
                           max_sims_per_move=a.sims, temps=temp_table(default_temp_scaling, 1, 42))
        for _ in range(24):                       # de-synchronise the games cheaply
            e.warmup_sims(8); e.play_moves(False)
        e.drain_samples(); e.drain_results()
        bts.append(torch.zeros(Bw, 4, 6, 7).pin_memory()); pts.append(torch.zeros(Bw, 7).pin_memory())
        vts.append(torch.zeros(Bw, 3).pin_memory()); evs.append(threading.Event())
        agents.append(SelfPlayAgent(i, _Game, ready, evs[i], bts[i], pts[i], vts[i], file_queue, _Sink(), completed, played,
                                    stop, pause, args, engine=e, stream_ordered=not a.e2e_sync))
    wrap = NNetWrapper(nnet=model, cuda=True, fused=(a.nn != "cudnn"), precision=a.precision if a.nn == "tc" else None)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = a.precision != "fp32"
    srv = torch.cuda.Stream(device=dev)
    from azb200.nnet import HostBatchServer
    server = HostBatchServer(wrap, stream=srv)
    if a.e2e_mode == "inline" and not a.e2e_sync:
        return run_e2e_inline(a, eng, agents, server, evs, ready, dev, world, Bw, old_tf32)
    for ag in agents:
        ag.start()

    def serve(until):
        with torch.cuda.stream(srv):
            while not until():
                try:
                    i = ready.get(timeout=0.5)
                except pyqueue.Empty:
                    continue
                if a.e2e_sync:
                    policy, value = wrap.process(bts[i])      # host -> device, network
                    pts[i].copy_(policy)                      # device -> host
                    vts[i].copy_(value)
                    evs[i].set()
                else:
                    # stream-ordered host tensors: same copies, ordered by CUDA events instead of host syncs
                    server.serve(agents[i], evs[i])

    serve(lambda: all(ag.batches >= a.sims + 2 for ag in agents))          # one untimed round per agent
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()          # stream-ordered mode: no backlog of enqueued work may leak into the timed region
    for ag in agents:
        ag.h2d_bytes = ag.d2h_bytes = 0
    n0 = sum(ag.batches for ag in agents)
    t0 = time.perf_counter()
    target = n0 + W * a.sims * a.e2e_steps
    serve(lambda: sum(ag.batches for ag in agents) >= target)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n1 = sum(ag.batches for ag in agents)
    stop.set()
    for ev in evs:
        ev.set()
    for ag in agents:
        ag.join(timeout=20)
    torch.backends.cudnn.allow_tf32 = old_tf32
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    n = torch.tensor([(n1 - n0) * Bw], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    steps = (n1 - n0) / float(W * a.sims)                 # move-rounds of all 8192 games
    obs_b, pv_b = eng.B * 4 * 6 * 7 * 4, eng.B * 10 * 4
    for ag in agents:
        ag.engine.close()
    return {"value": float(n.item()) / float(t.item()), "unit": UNIT,
            # per step: nnet.process uploads the observation batch, processBatch uploads policy/value
            "h2d_bytes_per_step": int(a.sims * obs_b + sum(ag.h2d_bytes for ag in agents) / steps),
            # per step: generateBatch downloads observations, the NN answers go to host tensors, samples are drained
            "d2h_bytes_per_step": int(a.sims * pv_b + sum(ag.d2h_bytes for ag in agents) / steps),
            "steps": steps, "agents": W, "games_per_agent": Bw,
            "file_queue": "one item per example" if a.e2e_item_queue else "azb200.coach.ExampleQueue (one block per move-round, every example kept)",
            "host_tensor_protocol": "host-synchronised" if a.e2e_sync else "stream-ordered (CUDA events)",
            "api": "Coach.processSelfPlayBatches loop: azb200.selfplay.SelfPlayAgent threads (generateBatch/processBatch/"
                   "playMoves) + NNetWrapper.process, pinned host tensors, ready queue + events"}


def run_e2e_inline(a, eng, agents, server, evs, ready, dev, world, Bw, old_tf32):
    """The same protocol driven by ONE host thread, the way the parity tests drive the reference in-process: per
    simulation and agent  generateBatch() -> HostBatchServer.serve() (NNetWrapper.process on the host batch tensor,
    answers into the host policy / value tensors) -> processBatch(), then playMoves().  Every call only enqueues:
    the host tensors are ordered by CUDA events, each agent has its own stream, so the copies and kernels of
    different agents overlap on the GPU while no Python thread ever waits for another."""
    import torch
    import torch.distributed as dist
    W = len(agents)
    for ag in agents:
        ag.stream = torch.cuda.Stream(device=dev)
        ag.step_graphs = False

    def one_round():
        for _ in range(a.sims):
            for i, ag in enumerate(agents):
                with torch.cuda.stream(ag.stream):
                    ag.generateBatch()
                j = ready.get_nowait()                       # id = self.ready_queue.get() (Coach.py:336)
                server.serve(agents[j], evs[j])
                with torch.cuda.stream(ag.stream):
                    ag.processBatch()
        for ag in agents:
            with torch.cuda.stream(ag.stream):
                ag.playMoves()

    def one_round_graphed():
        # the agents' own captured step graphs (SelfPlayAgent._graphed_round: [select, obs -> batch_tensor] /
        # [answers -> engine, expand/backup, select, obs -> batch_tensor] / [answers -> engine, expand/backup]),
        # interleaved by this one thread instead of one thread per agent: per simulation and agent the host issues two
        # graph launches and two event records
        for ag in agents:
            with torch.cuda.stream(ag.stream):
                if ag._g_first is None:
                    ag._prepare_graphs()
                ag._g_first.replay()
                ag._publish_batch()
        for s in range(a.sims):
            for _ in agents:
                j = ready.get_nowait()
                server.serve(agents[j], evs[j])
            for ag in agents:
                with torch.cuda.stream(ag.stream):
                    ag._await_answers()                  # batch_ready is already set: only the stream waits (answer_event)
                    if s + 1 < a.sims:
                        ag._g_mid.replay()
                        ag._publish_batch()
                    else:
                        ag._g_last.replay()
        for ag in agents:
            with torch.cuda.stream(ag.stream):
                ag.playMoves()

    def one_round_fused():
        # one graph per agent and move-round, the NN server's body captured inside (SelfPlayAgent.round_with_server):
        # the same copies through the same pinned host tensors, one host launch per agent and round
        for ag in agents:
            ag.round_with_server(server, a.sims)
        for ag in agents:
            with torch.cuda.stream(ag.stream):
                ag.playMoves()

    one_round()                                           # eager: one-time kernel / evaluator setup
    while not ready.empty():
        ready.get_nowait()
    if not a.e2e_eager:
        one_round = one_round_graphed if a.e2e_step_graphs else one_round_fused
    one_round(); one_round()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for ag in agents:
        ag.h2d_bytes = ag.d2h_bytes = 0
    n0 = sum(ag.batches for ag in agents)
    cpu0 = os.times()
    t0 = time.perf_counter()
    for _ in range(a.e2e_steps):
        one_round()
    t_issue = time.perf_counter() - t0                 # the host thread is done issuing; the GPU may still be working
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cpu1 = os.times()
    n1 = sum(ag.batches for ag in agents)
    torch.backends.cudnn.allow_tf32 = old_tf32
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    n = torch.tensor([(n1 - n0) * Bw], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    steps = (n1 - n0) / float(W * a.sims)
    obs_b, pv_b = eng.B * 4 * 6 * 7 * 4, eng.B * 10 * 4
    for ag in agents:
        ag.engine.close()
    return {"value": float(n.item()) / float(t.item()), "unit": UNIT,
            "h2d_bytes_per_step": int(a.sims * obs_b + sum(ag.h2d_bytes for ag in agents) / steps),
            "d2h_bytes_per_step": int(a.sims * pv_b + sum(ag.d2h_bytes for ag in agents) / steps),
            "steps": steps, "agents": W, "games_per_agent": Bw,
            # limiter diagnosis (this rank): share of the timed wall clock the issuing host thread was busy issuing, CPU
            # seconds of the process per wall second, and PCIe traffic each way
            "host_issue_fraction": t_issue / dt, "host_cpu_cores_busy": ((cpu1.user - cpu0.user) + (cpu1.system - cpu0.system)) / dt,
            "pcie_gbs_h2d": (a.sims * obs_b + sum(ag.h2d_bytes for ag in agents) / steps) * steps / dt / 1e9,
            "pcie_gbs_d2h": (a.sims * pv_b + sum(ag.d2h_bytes for ag in agents) / steps) * steps / dt / 1e9,
            "seconds": dt,
            "file_queue": "one item per example" if a.e2e_item_queue else "azb200.coach.ExampleQueue (one block per move-round, every example kept)",
            "host_tensor_protocol": "stream-ordered (CUDA events)",
            "issue": ("eager launches" if a.e2e_eager else "three step graphs per agent and simulation" if a.e2e_step_graphs
                      else "one CUDA graph per agent and move-round, NN-server body inside (SelfPlayAgent.round_with_server)"),
            "api": "Coach.processSelfPlayBatches data flow driven by one host thread: azb200.selfplay.SelfPlayAgent."
                   "generateBatch -> HostBatchServer.serve (NNetWrapper.process on the pinned host batch tensor) -> "
                   "processBatch, playMoves; every batch crosses PCIe in both directions"}


def gather_examples(kept, dev, rank, world):
    """BASELINE config 3: NCCL gather of the (s, pi, z) examples of the run to rank 0."""
    import torch
    import torch.distributed as dist
    from azb200.distributed import gather_examples_to_rank0
    obs, pi, z = kept
    gather_examples_to_rank0(obs[:1], pi[:1], z[:1])          # NCCL warm-up (communicator setup is not timed)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = gather_examples_to_rank0(obs, pi, z)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), (int(out[0].shape[0]) if rank == 0 else None)


if __name__ == "__main__":
    main()
