#!/usr/bin/env python
"""Where does the per-step time go at launch granularity?  Times, for C3 at B envs, CUDA graphs of 128 launches of
  (a) qg_observe only (records read + observation slab written: phase 2 alone),
  (b) qg_step without obs/mask (phase 1 alone + reward/done),
  (c) the full fused step,
  (d) a torch fill of the same slab (write ceiling at this launch size).
Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from qiskit_gym_b200 import BatchedEnv
from qiskit_gym_b200 import workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3_clifford8_full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
kind, n, gs, kw = W.baseline_configs()[cfg]
env = BatchedEnv(kind, n, gs, B, device=0, add_perms=False, **({} if kind == W.PAULI else {"add_inverts": False}), **kw)
env.set_state(W.random_targets(kind, n, gs, B, 1, scramble=64))
env.snapshot()
obs_size = int(np.prod(env.obs_shape())); A = env.num_actions()
nbuf = max(2, int(np.ceil(2 * 126e6 / (B * obs_size * 4))) + 1)
ring = torch.empty((nbuf, B, obs_size), dtype=torch.float32, device="cuda")
mring = torch.empty((nbuf, B, A), dtype=torch.bool, device="cuda")
T = 128
acts = torch.from_numpy(W.random_actions(np.random.default_rng(0), T, B, A)).cuda()
s = torch.cuda.Stream()


def graph(fn):
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        s.synchronize()
    return g


def timed(g, reps=10):
    with torch.cuda.stream(s):
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * T)   # us per launch


def f_obs():
    for t in range(T):
        env.observe(out=ring[t % nbuf])


def f_step_noobs():
    env.restore()
    for t in range(T):
        env.step(acts[t], obs=False, mask=False)


def f_step():
    env.restore()
    for t in range(T):
        env.step(acts[t], obs=ring[t % nbuf], mask=mring[t % nbuf])


def f_fill():
    for t in range(T):
        ring[t % nbuf].fill_(1.0)


out = {"config": cfg, "envs": B, "obs_bytes": B * obs_size * 4}
for name, fn in (("observe_only_us", f_obs), ("step_no_obs_us", f_step_noobs), ("full_step_us", f_step), ("torch_fill_obs_us", f_fill)):
    out[name] = timed(graph(fn))
out["observe_gbs"] = out["obs_bytes"] / out["observe_only_us"] / 1e3
out["fill_gbs"] = out["obs_bytes"] / out["torch_fill_obs_us"] / 1e3
print(json.dumps(out))
