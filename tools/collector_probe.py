#!/usr/bin/env python
"""The packed tensor-core collector alone (bench.py's `collector.tensor_core_packed` leg): time per decision, eager and CUDA graph.
    python tools/collector_probe.py [envs] [decisions]        (ncu --metrics gpu__time_duration.sum ... for the per-kernel list)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from qiskit_gym_b200 import BatchedEnv
from qiskit_gym_b200 import workloads as W
from qiskit_gym_b200.collector import RolloutCollector
from qiskit_gym_b200.search import BasicPolicy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
graph = (sys.argv[3] != "eager") if len(sys.argv) > 3 else True
kind, n, gateset, kw = W.baseline_configs()["C3_clifford8_full"]
env = BatchedEnv(kind, n, gateset, B, device=0, difficulty=64, depth_slope=2, max_depth=128, add_perms=False, add_inverts=False)
torch.manual_seed(0)
pol = BasicPolicy(env.obs_shape(), len(gateset), embedding_size=512, common_layers=(256,))
col = RolloutCollector(env, pol, use_twists=False, seed=0)
out = {"envs": B, "decisions": T}
for name, use_graph in (("eager", False), ("graph", True)):
    if name == "graph" and not graph:
        continue
    for _ in range(2):
        col.collect_packed(T, use_cuda_graph=use_graph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro = col.collect_packed(T, use_cuda_graph=use_graph)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[name] = {"us_per_decision": 1e3 * ms / T, "env_steps_per_s": B * T / (ms * 1e-3), "episodes": ro.episode_stats()[0]}
print(json.dumps(out))
