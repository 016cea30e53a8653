#!/usr/bin/env python
"""GPU probe of the tcgen05 leaf evaluator (csrc/azb_resnet_tc.cu): layer-by-layer
comparison with PyTorch fp32, final outputs, and timing against the mma.sync kernel.
  python scripts/tcprobe.py [--time]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

from test_fused_nn import _model, _obs
from azb200.fused_nn import FusedResNetEvaluator

dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def torch_layers(m, obs):
    """fp32 activations the kernel's epilogues produce, layer by layer: [B, H, W, C]."""
    out = []
    with torch.no_grad():
        x = F.relu(m.bn1(m.conv1(obs)))
        blocks = list(m.resnet)
        a = F.relu(blocks[0].bn1(x)) if blocks else x
        out.append(a)
        for i, blk in enumerate(blocks):
            b = F.relu(blk.bn2(blk.conv1(a)))
            out.append(b)
            x = x + blk.conv2(b)
            a = F.relu(blocks[i + 1].bn1(x)) if i + 1 < len(blocks) else x
            out.append(a)
    return [t.permute(0, 2, 3, 1).contiguous() for t in out]


def check(depth, batch):
    m = _model(depth=depth).to(dev)
    obs = _obs(batch).to(dev)
    pol = torch.zeros(batch, 7, device=dev); val = torch.zeros(batch, 3, device=dev)
    ev = FusedResNetEvaluator(m, obs, pol, val, kernel="tc")
    want = torch_layers(m, obs)
    for l in range(1 + 2 * depth):
        got = ev.debug_layer(l)
        torch.cuda.synchronize()
        err = (got - want[l]).abs()
        scale = want[l].abs().max().item()
        print(f"depth {depth} batch {batch} layer {l}: max err {err.max().item():.4g} mean {err.mean().item():.4g} (|act| max {scale:.3g})",
              flush=True)
    ev()
    torch.cuda.synchronize()
    with torch.no_grad():
        lp, lv = m(obs)
    ep, evl = (pol - lp.exp()).abs().max().item(), (val - lv.exp()).abs().max().item()
    print(f"depth {depth} batch {batch}: policy err {ep:.4g} value err {evl:.4g} sum {pol.sum(1).mean().item():.6f}", flush=True)
    return ep, evl


def timeit(kernel, batch=8192, reps=50):
    m = _model(depth=4).to(dev)
    obs = _obs(batch).to(dev)
    pol = torch.zeros(batch, 7, device=dev); val = torch.zeros(batch, 3, device=dev)
    ev = FusedResNetEvaluator(m, obs, pol, val, kernel=kernel)
    for _ in range(5):
        ev()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ev()
    e1.record()
    torch.cuda.synchronize()
    us = 1000.0 * e0.elapsed_time(e1) / reps
    print(f"{kernel}: {us:.1f} us per {batch} boards = {batch / us:.2f} M evals/s, {batch * 8.1e6 / us / 1e6:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if "--time-only" in sys.argv:
        timeit("tc", reps=3)
        sys.exit(0)
    for depth, batch in ((0, 16), (1, 16), (4, 40), (4, 1000)):
        check(depth, batch)
    if "--time" in sys.argv:
        timeit("mma")
        timeit("tc")
