"""Pipeline trace of CTA 0 of the trunk kernel (DBG build, azb_nng_trace): per (layer, tile) when the MMAs start to issue,
are committed, when the epilogue sees the accumulator and when it hands the tile on.
python scripts/nn_trace.py [geom] [precision] [batch]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-general_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from azb200 import nn_tc  # noqa: E402
from test_nn_tc import GEOMS, _model, _obs  # noqa: E402

geom = sys.argv[1] if len(sys.argv) > 1 else "brandubh"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x2"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 3915
dev = torch.device("cuda")
m = _model(geom).to(dev)
obs = _obs(geom, batch).to(dev)
pol = torch.zeros(batch, GEOMS[geom]["A"], device=dev); val = torch.zeros(batch, 3, device=dev)
ev = nn_tc.TensorCoreEvaluator(m, obs, pol, val, precision=prec)
layer = int(sys.argv[4]) if len(sys.argv) > 4 else 2 * ev.net.depth
ev.debug_layer(layer); ev.debug_layer(layer)
torch.cuda.synchronize()
buf = (C.c_longlong * 1024)()
assert ev.lib.azb_nng_trace(buf, 1024) == 0
lay = ev.layout
tiles = max(1, min(lay["boards_per_cta"] // 2, 64))
layers = 1 + 2 * ev.net.depth
print(f"{geom} {prec} pair={ev.pair}: cycles relative to the first issue; columns: issue start, committed, epilogue start, handed on")
t0 = buf[0]
print(f"set-up detail: frame zeroed by thread 0 {buf[1008] - t0}, barriers initialised {buf[1009] - t0}, init fence done {buf[1010] - t0}")
print(f"set-up detail: alloc starts {buf[1004] - t0}, alloc done {buf[1005] - t0}, frame zeroed {buf[1006] - t0}, after __syncthreads {buf[1007] - t0}")
print(f"CTA entry {buf[1000] - t0}, set-up done {buf[1001] - t0}, observations staged {buf[1002] - t0}, first issue 0, exit {buf[1003] - t0}")
plan = (C.c_int32 * 6)()
sms = torch.cuda.get_device_properties(0).multi_processor_count
assert ev.lib.azb_nng_tile_plan(ev.net.channels, batch, sms, 1 if ev.pair else 0, int(os.environ.get("AZB_NNG_PERSIST", "1")), 0, plan) == 0
units, n_tiles, first, rounds, tiles, big = plan
print(f"unit 0 of {units}: {n_tiles} tiles in {rounds} rounds ({big} of {tiles}, the others of {tiles - 1}); {layers} layers")
ev_rows = []
g = 0
n_entries = layers * (big * tiles + (rounds - big) * (tiles - 1)) if big < rounds else rounds * layers * tiles
while g < 256 and buf[4 * g] != 0 and g < n_entries:
    a, b, c, d = (buf[4 * g + i] - t0 for i in range(4))
    ev_rows.append((g, a, b, c, d))
    if os.environ.get("NN_TRACE_FULL"):
        print(f"g {g:3d}: issue {a:8d} commit {b:8d} (+{b - a:5d})  epi start {c:8d} (+{c - b:5d} after commit)  done {d:8d} (epilogue {d - c:5d})")
    g += 1
# summary: per (round, layer) the span of its tiles' MMAs, and the idle time of the issue stream before each tile
print("round layer: first issue, last commit, span, issue time (sum of commit - issue), idle before its tiles (sum of issue - previous commit)")
prev_commit = 0
pos = 0
for v in range(rounds * layers):
    tc = tiles if (v // layers) < big else tiles - 1
    rows = ev_rows[pos:pos + tc]
    pos += tc
    if not rows:
        break
    busy = sum(b - a for _, a, b, _, _ in rows)
    idle = 0
    for _, a, b, _, _ in rows:
        idle += max(0, a - prev_commit)
        prev_commit = b
    epi = [d - c for _, _, _, c, d in rows]
    lag = [c - b for _, _, b, c, _ in rows]
    print(f"r {v // layers} l {v % layers}: {rows[0][1]:8d} {rows[-1][2]:8d} span {rows[-1][2] - rows[0][1]:6d} issue {busy:6d} idle {idle:6d}  "
          f"epilogue mean {sum(epi) // len(epi):5d}  commit->epilogue start mean {sum(lag) // len(lag):5d}")
if ev_rows:
    print(f"last commit {ev_rows[-1][2]}, last epilogue done {max(r[4] for r in ev_rows)}, exit {buf[1003] - t0}")

# every CTA: life time and the gap to the next CTA on the same SM
import collections
n = 2048
cb = (C.c_longlong * (3 * n))()
assert ev.lib.azb_nng_cta_trace(cb, n) == 0
per_sm = collections.defaultdict(list)
for i in range(n):
    sm, a, b = cb[3 * i], cb[3 * i + 1], cb[3 * i + 2]
    if b > a > 0:
        per_sm[sm].append((a, b, i))
t_first = min(v[0][0] for v in per_sm.values() if v)
lives, gaps, ends = [], [], []
for sm, v in per_sm.items():
    v.sort()
    for k, (a, b, i) in enumerate(v):
        lives.append(b - a)
        if k:
            gaps.append(a - v[k - 1][1])
    ends.append(v[-1][1] - t_first)
lives.sort(); gaps.sort()
q = lambda x, f: x[min(len(x) - 1, int(f * len(x)))] if x else None
print(f"{sum(len(v) for v in per_sm.values())} CTAs on {len(per_sm)} SMs: life ns p10/p50/p90 = {q(lives, .1)}/{q(lives, .5)}/{q(lives, .9)}, "
      f"gap to the next CTA on the SM p10/p50/p90 = {q(gaps, .1)}/{q(gaps, .5)}/{q(gaps, .9)}, SMs finish at {min(ends)} .. {max(ends)} ns")
sm0 = sorted(per_sm)[0]
print("SM", sm0, [(a - t_first, b - t_first, i) for a, b, i in per_sm[sm0]])
