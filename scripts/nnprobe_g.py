"""Time the leaf evaluators on one GPU (CUDA events, L2 flushed between iterations is not needed: obs of 8192 boards
is 5.5 MB and the weights are meant to stay in L2 -- this is the in-loop situation).  Usage:
python scripts/nnprobe_g.py [--game connect4|brandubh] [--batch N]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from azb200 import nnet as aznet, nn_tc  # noqa: E402
from azb200.fused_nn import FusedResNetEvaluator, supported_tc  # noqa: E402


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--game", default="connect4")
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    from test_nn_tc import GEOMS, _model, _obs, _want
    geom = a.game
    batch = a.batch or (8192 if geom.startswith("connect4") else 4096)
    dev = torch.device("cuda")
    m = _model(geom).to(dev)
    obs = _obs(geom, batch).to(dev)
    A = GEOMS[geom]["A"]
    want = _want(m, obs)
    flops = {"connect4": 8.084e6, "brandubh": 34.0e6}.get(geom, 0.0)
    for prec in ("bf16x2", "fp16", "bf16"):
        pol = torch.zeros(batch, A, device=dev); val = torch.zeros(batch, 3, device=dev)
        ev = nn_tc.TensorCoreEvaluator(m, obs, pol, val, precision=prec)
        us = timeit(ev)
        err = max((pol - want[0]).abs().max().item(), (val - want[1]).abs().max().item())
        print(f"{geom} B={batch} tcg/{prec:7s}: {us:8.1f} us  {batch / us:7.2f} M evals/s  {flops * batch / us * 1e-6:7.1f} TFLOP/s(useful)  max err {err:.2e}",
              flush=True)
    if supported_tc(m):
        pol = torch.zeros(batch, A, device=dev); val = torch.zeros(batch, 3, device=dev)
        ev = FusedResNetEvaluator(m, obs, pol, val, kernel="tc")
        us = timeit(ev)
        err = max((pol - want[0]).abs().max().item(), (val - want[1]).abs().max().item())
        print(f"{geom} B={batch} r1 tc/bf16  : {us:8.1f} us  {batch / us:7.2f} M evals/s  max err {err:.2e}", flush=True)
    for prec in ("tf32", "fp32"):
        pol = torch.zeros(batch, A, device=dev); val = torch.zeros(batch, 3, device=dev)
        ev = aznet.LeafEvaluator(m, obs, pol, val, precision=prec)
        us = timeit(ev, iters=10, warm=2)
        err = max((pol - want[0]).abs().max().item(), (val - want[1]).abs().max().item())
        print(f"{geom} B={batch} cudnn/{prec:5s}: {us:8.1f} us  {batch / us:7.2f} M evals/s  max err {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
