#!/usr/bin/env python
"""Which output stream costs what in the replay launch (C3 Clifford 8q by default): times qg_replay with subsets of
{obs, mask, reward, done, success} and a few batch sizes.  One JSON line per variant.
    python tools/replay_ablation.py [config] [envs ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_gym_b200 import BatchedEnv, workloads as W  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C3_clifford8_full"
    sizes = [int(x) for x in sys.argv[2:]] or [65536]
    kind, n, gateset, kw = W.baseline_configs()[cfg]
    T, RING = 128, int(os.environ.get("RING", 5))
    for B in sizes:
        env = BatchedEnv(kind, n, gateset, B, device=0, add_inverts=False, add_perms=False, **kw)
        env.set_state(W.random_targets(kind, n, gateset, min(B, 65536), seed=1)[np.arange(B) % min(B, 65536)])
        env.snapshot()
        A, O = env.num_actions(), int(np.prod(env.obs_shape()))
        dev = env.device
        actions = torch.randint(0, A, (T, B), dtype=torch.int32, device=dev)
        obs = torch.empty((RING, B, O), dtype=torch.float32, device=dev)
        mask = torch.empty((RING, B, A), dtype=torch.bool, device=dev)
        rew = torch.empty((T, B), dtype=torch.float32, device=dev)
        done = torch.empty((T, B), dtype=torch.bool, device=dev)
        suc = torch.empty((T, B), dtype=torch.bool, device=dev)
        variants = {
            "full": dict(obs=obs, mask=mask, reward=rew, done=done, success=suc),
            "obs+mask": dict(obs=obs, mask=mask),
            "obs+rds": dict(obs=obs, reward=rew, done=done, success=suc),
            "obs": dict(obs=obs),
            "mask+rds": dict(mask=mask, reward=rew, done=done, success=suc),
            "none": dict(),
        }
        for name, kwv in variants.items():
            def run():
                env.restore()
                env.replay(actions, **kwv)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nbytes = T * B * ((4 * O if "obs" in kwv else 0) + (A if "mask" in kwv else 0) + (6 if "reward" in kwv else 0) + 4)
            print(json.dumps({"config": cfg, "envs": B, "variant": name, "ms": ms, "env_steps_per_s": T * B / ms * 1e3, "stream_gbs": nbytes / ms / 1e6,
                              "stagger": os.environ.get("QG_STAGGER_NS", "default"), "ring": RING}), flush=True)
        del env, obs, mask


if __name__ == "__main__":
    main()
