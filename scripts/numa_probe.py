#!/usr/bin/env python
"""Does the NUMA node the pinned host buffers live on explain the slow host->device direction?  For every node: pin the
thread to the node's cores, allocate + touch pinned buffers there, time H2D / D2H at the e2e leg's batch size."""
import glob
import os
import subprocess
import time

import torch


def cpus(path):
    out = []
    for part in open(path).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def main():
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
    except Exception as ex:
        print("nvidia-smi topo failed:", ex)
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    print("numa nodes:", [(os.path.basename(n), len(cpus(n + "/cpulist"))) for n in nodes], "allowed cpus:", len(os.sched_getaffinity(0)))
    dev = torch.device("cuda", 0)
    torch.zeros(1, device=dev)
    allowed = os.sched_getaffinity(0)
    cases = [("default", allowed)] + [(os.path.basename(n), set(cpus(n + "/cpulist")) & allowed) for n in nodes]
    for name, cs in cases:
        if not cs:
            continue
        os.sched_setaffinity(0, cs)
        for mb in (5.5, 64):
            n = int(mb * 1e6 / 4)
            h = torch.empty(n, pin_memory=True); h.fill_(1.0)
            d = torch.zeros(n, device=dev)
            res = {}
            for what, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
                for _ in range(5):
                    f()
                torch.cuda.synchronize(); t0 = time.perf_counter()
                for _ in range(50):
                    f()
                torch.cuda.synchronize(); res[what] = mb * 1e6 / ((time.perf_counter() - t0) / 50) / 1e9
            print(f"{name:8s} ({len(cs):3d} cpus) {mb:5.1f} MB: h2d {res['h2d']:5.1f} GB/s  d2h {res['d2h']:5.1f} GB/s", flush=True)
        os.sched_setaffinity(0, allowed)


if __name__ == "__main__":
    main()
