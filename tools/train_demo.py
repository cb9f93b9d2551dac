#!/usr/bin/env python
"""Short PPO runs on the engine (RLSynthesis.learn): prints one line per iteration.
    python tools/train_demo.py [perm4|lf4|cliff3] [iterations]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_gym_b200.specs import SynthSpec  # noqa: E402
from qiskit_gym_b200.rl import RLSynthesis  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:                                   # torchrun: one process per GPU, gradients averaged over NCCL (ppo.sync_gradients)
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[1] if len(sys.argv) > 1 else "perm4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
line4 = [(0, 1), (1, 2), (2, 3)]
both = line4 + [(b, a) for a, b in line4]
if which == "perm4":
    env = SynthSpec.from_coupling_map("PermutationEnv", line4, difficulty=1, depth_slope=2, max_depth=32)
elif which == "lf4":
    env = SynthSpec.from_coupling_map("LinearFunctionEnv", both, basis_gates=("CX",), difficulty=1, depth_slope=2, max_depth=64)
else:
    tri = [(0, 1), (1, 0), (1, 2), (2, 1)]
    env = SynthSpec.from_coupling_map("CliffordEnv", tri, basis_gates=("H", "S", "CX"), difficulty=1, depth_slope=2, max_depth=64)
torch.manual_seed(0)
cfg = {"collecting": {"num_episodes": int(os.environ.get("EPISODES", 512))}, "training": {"num_epochs": 4}, "optimizer": {"lr": float(os.environ.get("LR", 2e-3))},
       "learning": {"diff_max": 32}, "evals": {"ppo_deterministic": {"num_episodes": 128}, "ppo_10": {"num_episodes": 128, "deterministic": False, "num_searches": 10}}}
algo = os.environ.get("ALGO", "PPO")
if algo == "AZ":
    cfg["collecting"].update({"num_episodes": int(os.environ.get("EPISODES", 256)), "num_mcts_searches": int(os.environ.get("SIMS", 16)), "C": 1.41})
    cfg["learning"]["diff_metric"] = "mcts"
    cfg["evals"] = {"ppo_deterministic": {"num_episodes": 64}, "mcts": {"num_episodes": 64, "num_mcts_searches": int(os.environ.get("SIMS", 16))}}
rls = RLSynthesis(env, cfg, {"embedding_size": 64, "common_layers": [64]}, device=local, algorithm_cls=f"twisterl.rl.{algo}")
t0 = time.time()
rls.learn(num_iterations=iters, log=(lambda r: print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)) if rank == 0 else None)
if world > 1:
    w = torch.cat([p.detach().reshape(-1) for p in rls.policy.parameters()])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    if rank == 0:
        print("replicas identical after training:", all(torch.equal(ws[0], x) for x in ws))
    dist.destroy_process_group()
if rank == 0:
    print("total s", round(time.time() - t0, 2))
