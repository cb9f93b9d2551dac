#!/usr/bin/env python
"""One replay launch without observation / mask outputs (the step logic alone: phase 1 of k_step), for an ncu source-level
capture:  ncu --set full --import-source on -k regex:k_step -c 1 -o gpurun_out/phase1 python tools/phase1_profile.py [config]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_gym_b200 import BatchedEnv, workloads as W  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3_clifford8_full"
kind, n, gateset, kw = W.baseline_configs()[cfg]
B, T = 65536, 128
env = BatchedEnv(kind, n, gateset, B, device=0, add_inverts=False, add_perms=False, **kw)
env.set_state(W.random_targets(kind, n, gateset, 4096, seed=1)[np.arange(B) % 4096])
actions = torch.randint(0, env.num_actions(), (T, B), dtype=torch.int32, device=env.device)
rew = torch.empty((T, B), dtype=torch.float32, device=env.device)
env.replay(actions, reward=rew)
torch.cuda.synchronize()
print("ok", float(rew.sum()))
