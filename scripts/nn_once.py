"""One evaluation of the tcgen05 evaluator (for ncu).  python scripts/nn_once.py [connect4|brandubh] [precision] [batch] [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from azb200 import nn_tc  # noqa: E402
from test_nn_tc import GEOMS, _model, _obs  # noqa: E402

geom = sys.argv[1] if len(sys.argv) > 1 else "connect4"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x2"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else (3915 if geom.startswith("brandubh") else 6960)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda")
m = _model(geom).to(dev)
obs = _obs(geom, batch).to(dev)
A = GEOMS[geom]["A"]
pol = torch.zeros(batch, A, device=dev); val = torch.zeros(batch, 3, device=dev)
ev = nn_tc.TensorCoreEvaluator(m, obs, pol, val, precision=prec)
ev()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ev()
e1.record()
torch.cuda.synchronize()
print(f"ok {geom} {ev.precision} batch {batch}: {e0.elapsed_time(e1) * 1000 / iters:.1f} us per evaluation (trunk + head), sum {float(pol.sum()):.3f}")
