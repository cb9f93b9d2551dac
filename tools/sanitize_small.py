"""A small replay (warp-pair kernels), a staged host-packed replay and a tensor-core policy forward for compute-sanitizer:
    compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_gym_b200 import BatchedEnv  # noqa: E402
from qiskit_gym_b200 import workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
for name, B, T in (("C3_clifford8_full", 96, 12), ("C1_perm_grid3", 70, 10), ("C4_pauli10_line", 64, 9)):
    kind, n, gateset, kw = W.baseline_configs()[name]
    pk = dict(kw)
    if kind != W.PAULI:
        pk["add_inverts"] = False
    env = BatchedEnv(kind, n, gateset, B, device=0, max_depth=T, add_perms=False, **pk)
    env.set_state(W.random_targets(kind, n, gateset, B, 3, scramble=16))
    A = len(gateset)
    rng = np.random.Generator(np.random.PCG64(5))
    acts = W.random_actions(rng, T, B, A)
    osz = int(np.prod(env.obs_shape()))
    obs = torch.empty((T, B, osz), dtype=torch.float32, device=dev)
    mask = torch.empty((T, B, A), dtype=torch.bool, device=dev)
    rew = torch.empty((T, B), dtype=torch.float32, device=dev)
    env.snapshot()
    env.replay(torch.from_numpy(acts).to(dev), obs=obs, mask=mask, reward=rew)
    torch.cuda.synchronize()
    print(name, "replay ok", float(rew.sum()))
# staged host-packed path (>= 1 MB of actions)
kind, n, gateset, kw = W.baseline_configs()["C2_lf8_line"]
B, T = 16391, 70
env = BatchedEnv(kind, n, gateset, B, device=0, max_depth=T, add_perms=False, add_inverts=False)
env.set_state(W.random_targets(kind, n, gateset, B, 3, scramble=16))
rng = np.random.Generator(np.random.PCG64(6))
h_a = env.host_buffer((T, B), np.uint8); h_a[:] = W.random_actions(rng, T, B, len(gateset)).astype(np.uint8)
h_r = env.host_buffer((T, B), np.float32); h_d = env.host_buffer((env.flag_words(), T), np.uint32)
env.replay_host_packed(h_a, h_d, None, reward=h_r)
print("host packed ok", float(h_r.sum()))
