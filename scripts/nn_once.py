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
batch = int(sys.argv[3]) if len(sys.argv) > 3 else (6960 if geom == "connect4" else 3915)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda")
m = _model(geom).to(dev)
obs = _obs(geom, batch).to(dev)
A = GEOMS[geom]["A"]
pol = torch.zeros(batch, A, device=dev); val = torch.zeros(batch, 3, device=dev)
ev = nn_tc.TensorCoreEvaluator(m, obs, pol, val, precision=prec)
for _ in range(iters):
    ev()
torch.cuda.synchronize()
print("ok", float(pol.sum()))
