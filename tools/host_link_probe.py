#!/usr/bin/env python
"""What the host link of this box sustains when every GPU talks to host memory at once: plain pinned cudaMemcpyAsync D2H, H2D and both,
all ranks together (torchrun, one rank per GPU), aggregate GB/s.  The yardstick for bench.py's end-to-end numbers at 2 / 4 / 8 GPUs:
qg_replay_host_packed moves 1 B (actions, H2D) + 4.25 B (reward + flag bits, D2H) per env-step per GPU.
    python -m torch.distributed.run --nproc-per-node N tools/host_link_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/qg_probe_nccl.%h.%p.log")
    dist.init_process_group("nccl", device_id=dev)
MB = 256
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(); h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB << 20, dtype=torch.uint8, device=dev); d_out = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, reps=8):
    def once():
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s1):
                h_out.copy_(d_out, non_blocking=True)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d_in.copy_(h_in, non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    per_dir = reps * (MB << 20) / sec / 1e9
    return per_dir


out = {"n_gpus": world, "buffer_mb": MB}
for mode in ("d2h", "h2d", "both"):
    g = run(mode)
    out[mode + "_gbs_per_gpu_per_direction"] = g
    out[mode + "_gbs_aggregate"] = g * world * (2 if mode == "both" else 1)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
