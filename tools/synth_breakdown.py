#!/usr/bin/env python
"""Where a synth-search decision spends its time (C5: PermutationGym 27q heavy-hex, 1000 rollouts): CUDA-event timings of
the policy forward and of the fused sample+step kernel, each alone, for the fused (packed-bit) and the PyTorch backend."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qiskit_gym_b200 import BatchedEnv, workloads as W
from qiskit_gym_b200.policy import FusedPolicy
from qiskit_gym_b200.search import BasicPolicy


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                fn()
    return lambda: g.replay()


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    kind, n, gs, kw = W.baseline_configs()["C5_perm27_heavyhex"]
    env = BatchedEnv(kind, n, gs, R, max_depth=1 << 20, add_inverts=False, add_perms=False)
    rng = np.random.Generator(np.random.PCG64(1))
    env.set_state(rng.permutation(n).astype(np.int64).tolist())
    env.search_begin(0, 0)
    torch.manual_seed(0)
    pol = BasicPolicy([n, n], len(gs), embedding_size=512, common_layers=(256,)).cuda().eval()
    fused = FusedPolicy(pol)
    bits = env.new_obs_bits()
    env.observe_bits(bits)
    env.observe()
    probs = torch.full((R, len(gs)), 1.0 / len(gs), dtype=torch.float32, device=env.device)
    out = {"rollouts": R}

    def torch_policy():
        with torch.no_grad():
            lg, _ = pol(env.obs)
            torch.softmax(lg, dim=-1, out=probs)

    out["fused_policy_us"] = timed(graphed(lambda: fused.forward_bits(bits, probs=probs))) / 20
    out["torch_policy_us"] = timed(graphed(torch_policy)) / 20
    out["search_step_bits_us"] = timed(graphed(lambda: env.search_step_bits(probs, bits))) / 20
    out["search_step_dense_us"] = timed(graphed(lambda: env.search_step(probs, obs=True))) / 20
    out["fused_decision_us"] = timed(graphed(lambda: (fused.forward_bits(bits, probs=probs), env.search_step_bits(probs, bits)))) / 20
    out["torch_decision_us"] = timed(graphed(lambda: (torch_policy(), env.search_step(probs, obs=True)))) / 20
    print(json.dumps(out))


if __name__ == "__main__":
    main()
