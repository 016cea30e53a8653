#!/usr/bin/env python
"""Pinned host<->device copy bandwidth at the e2e leg's batch sizes (one direction, both directions at once)."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "alphazero-general_b200"))
from azb200 import _capi
lib = _capi.load()
lib.azb_upload_pinned.restype = C.c_int
lib.azb_upload_pinned.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
dev = torch.device("cuda")
for mb in (2.75, 5.5, 64):
    n = int(mb * 1e6 / 4)
    h1, h2 = torch.zeros(n).pin_memory(), torch.zeros(n).pin_memory()
    d1, d2 = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(f, reps=50):
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        s1.synchronize(); s2.synchronize()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1000
    def h2d():
        with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    def both():
        h2d(); d2h()
    def up():
        assert lib.azb_upload_pinned(d1.data_ptr(), h1.data_ptr(), n * 4, C.c_void_p(s1.cuda_stream)) == 0
    def up_both():
        up(); d2h()
    import time
    res = {}
    h1.copy_(torch.arange(n) % 977)
    up(); torch.cuda.synchronize()
    assert torch.equal(d1.cpu(), h1), "upload kernel mismatch"
    for name, f in (("h2d", h2d), ("d2h", d2h), ("both", both), ("up", up), ("up_both", up_both)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(100): f()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 100
        res[name] = dt * 1e6
    print(f"{mb} MB: h2d {res['h2d']:.0f} us ({mb*1e6/res['h2d']/1e3:.1f} GB/s), d2h {res['d2h']:.0f} us ({mb*1e6/res['d2h']/1e3:.1f} GB/s), "
          f"both directions at once {res['both']:.0f} us; SM upload {res['up']:.0f} us ({mb*1e6/res['up']/1e3:.1f} GB/s), with d2h {res['up_both']:.0f} us")
