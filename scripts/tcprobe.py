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


def phases(batch=8192, code=99, show=True):
    """Per-CTA / per-warp phase clocks from the kernel's timing hook (dump_layer = 99)."""
    import ctypes as C
    m = _model(depth=4).to(dev)
    obs = _obs(batch).to(dev)
    pol = torch.zeros(batch, 7, device=dev); val = torch.zeros(batch, 3, device=dev)
    ev = FusedResNetEvaluator(m, obs, pol, val, kernel="tc")
    ctas = -(-batch // 16)
    buf = torch.zeros(ctas * 16 * 56 * 32, device=dev)
    for _ in range(3):
        ev()
    rc = ev.lib.azb_nn_forward_tc_debug(C.byref(ev.w), obs.data_ptr(), pol.data_ptr(), val.data_ptr(), batch,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream), buf.data_ptr(), code)
    torch.cuda.synchronize()
    t = buf[:ctas * 32 * 8].view(ctas, 32, 8).cpu()
    nw = 28
    if not show:
        print(f"dbg {99 - code}: trunk {t[:, :nw, 1].mean().item():.0f} cycles, CTA total {t[:, 0, :4].sum(1).mean().item():.0f}")
        return
    names = ["prologue", "trunk(own)", "wait-slowest", "heads"]
    for i, n in enumerate(names):
        v = t[:, :nw, i]
        print(f"{n:14s} mean {v.mean().item():9.0f} max {v.max().item():9.0f} | issuer {t[:, 0:3, i].mean().item():9.0f} producer {t[:, 3, i].mean().item():9.0f} "
              f"epilogue {t[:, 4:nw, i].mean().item():9.0f}")
    tl = buf[1024 * 256: 1024 * 256 + 63 * 16].view(63, 16).cpu()
    print(" g  l t | issuer: start  rows-ok ring-ok turn-ok committed | epilogue: poll-start full   released(after bar) | done: first..last warp")
    for g in range(14, 44):
        r = tl[g].tolist()
        print(f"{g:2d} {g // 7:2d} {g % 7} | {r[0]:7.0f} {r[1]:7.0f} {r[2]:7.0f} {r[3]:7.0f} {r[4]:7.0f} | {r[5]:7.0f} {r[6]:7.0f} {r[7]:7.0f} | {min(r[8:16]):7.0f} {max(r[8:16]):7.0f}")
    tot = t[:, 0, :4].sum(1)
    print(f"CTA total cycles: mean {tot.mean().item():.0f} min {tot.min().item():.0f} max {tot.max().item():.0f}")
    sm = t[:, 0, 4].long()
    per_sm = torch.bincount(sm, minlength=148)
    print("CTAs per SM: min", per_sm.min().item(), "max", per_sm.max().item(), "SMs used", int((per_sm > 0).sum()))


if __name__ == "__main__":
    if "--phases" in sys.argv:
        phases(code=int(sys.argv[sys.argv.index("--phases") + 1]) if sys.argv[-1] != "--phases" else 99)
        sys.exit(0)
    if "--time-only" in sys.argv:
        timeit("tc", reps=3)
        sys.exit(0)
    for depth, batch in ((0, 16), (1, 16), (4, 40), (4, 1000)):
        check(depth, batch)
    if "--time" in sys.argv:
        timeit("mma")
        timeit("tc")
