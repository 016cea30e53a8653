#!/usr/bin/env python
"""H2D / D2H bandwidth from write-combined pinned host memory (cudaHostAllocWriteCombined) vs ordinary pinned memory at
the e2e leg's batch sizes: the host batch tensors are only ever touched by DMA (device->host by the agent, host->device
by the NN server), so a write-combined allocation would be legal there if it helped."""
import ctypes as C
import time

import numpy as np
import torch

rt = C.CDLL("libcudart.so")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)


def host_tensor(n, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), n * 4, flags) == 0
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,))
    return torch.from_numpy(a)


for mb in (2.75, 5.5, 64):
    n = int(mb * 1e6 / 4)
    d = torch.zeros(n, device=dev)
    for name, flags in (("pinned", 0), ("write-combined", 4), ("torch pin_memory", None)):
        h = torch.empty(n, pin_memory=True) if flags is None else host_tensor(n, flags)
        d.fill_(1.0)
        h.copy_(d, non_blocking=True); torch.cuda.synchronize()
        res = {}
        for what, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
            for _ in range(5):
                f()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(50):
                f()
            torch.cuda.synchronize(); res[what] = mb * 1e6 / ((time.perf_counter() - t0) / 50) / 1e9
        print(f"{mb:5.2f} MB {name:17s} is_pinned={h.is_pinned()}  h2d {res['h2d']:5.1f} GB/s  d2h {res['d2h']:5.1f} GB/s", flush=True)
