#!/usr/bin/env python
"""Write-bandwidth ceilings on this GPU, to put the step kernel's roofline fraction in context
(MEASURED_PEAKS.json's hbm_gbs is a copy: read + write bytes).  Prints one JSON line.
  - fill_1g:   torch fill of a 1 GiB float buffer (write-only stream, one long launch)
  - fill_82m:  128 back-to-back fills of one 82 MB slab out of a 5-slab ring inside a CUDA graph: the size
               of one single-step launch of C3 at 65 536 envs, i.e. the launch-granularity ceiling
  - copy_1g:   the MEASURED_PEAKS method (b.copy_(a), read+write bytes)"""
import json

import torch


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    dev = torch.device("cuda", 0)
    out = {}
    big = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    for _ in range(3):
        big.fill_(1.0)
    out["fill_1g_gbs"] = big.numel() * 4 / min(timed(lambda: big.fill_(1.0), 5) for _ in range(3)) / 1e9
    src = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    for _ in range(3):
        big.copy_(src)
    out["copy_1g_gbs"] = 2 * big.numel() * 4 / min(timed(lambda: big.copy_(src), 5) for _ in range(3)) / 1e9
    del src
    n = 65536 * 1249 // 4
    ring = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(5)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for r in ring:
            r.fill_(0.0)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for t in range(128):
                ring[t % 5].fill_(1.0)
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10):
            g.replay()
        e1.record(s)
        s.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3 / (10 * 128)
    out["fill_82m_graph_gbs"] = n * 4 / dt / 1e9
    out["fill_82m_graph_us"] = dt * 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
