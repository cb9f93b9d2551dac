"""Where does the end-to-end episode (qg_restore + qg_replay_host_packed + synchronise) spend its time?

Times, for BASELINE config C3 at 65 536 envs x 128 env-steps: the replay kernel alone with device buffers (qg_replay_packed) and with
pinned host buffers (kernel-side, CUDA events around the launch only), the restore, and the host wall clock of the whole call sequence.
usage: python tools/e2e_probe.py [config] > gpurun_out/<tag>_e2e_probe.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qiskit_gym_b200 import BatchedEnv  # noqa: E402
from qiskit_gym_b200 import workloads as W  # noqa: E402
from qiskit_gym_b200._lib import lib  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3_clifford8_full"
    kind, n, gateset, kw = W.baseline_configs()[name]
    B, T = 65536, 128
    A = len(gateset)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib().qg_bind_thread_to_device(0)
    pk = dict(kw)
    if kind != W.PAULI:
        pk["add_inverts"] = False
    env = BatchedEnv(kind, n, gateset, B, device=0, max_depth=T, add_perms=False, **pk)
    obs_size = int(np.prod(env.obs_shape()))
    env.set_state(W.random_targets(kind, n, gateset, B, 7, scramble=256))
    env.snapshot()
    rng = np.random.Generator(np.random.PCG64(7))
    actions_h = W.random_actions(rng, T, B, A).astype(np.uint8)
    nbuf = T          # one observation slab per step, as in bench.py
    obs_ring = torch.empty((nbuf, B, obs_size), dtype=torch.float32, device=dev)
    mask_ring = torch.empty((nbuf, B, A), dtype=torch.bool, device=dev)
    tiles = env.flag_words()
    h_a8 = env.host_buffer((T, B), np.uint8); h_a8[:] = actions_h
    h_rw = env.host_buffer((T, B), np.float32)
    h_db = env.host_buffer((tiles, T), np.uint32); h_sb = env.host_buffer((tiles, T), np.uint32)
    d_a8 = torch.from_numpy(actions_h).to(dev)
    d_rw = torch.empty((T, B), dtype=torch.float32, device=dev)
    d_db = torch.empty((tiles, T), dtype=torch.int32, device=dev); d_sb = torch.empty((tiles, T), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    out = {"config": name, "envs": B, "steps": T}

    def events(fn, reps=10, pre=None):
        ts = []
        with torch.cuda.stream(stream):
            for _ in range(reps + 2):
                if pre:
                    pre()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); fn(); e1.record(stream); stream.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
        return float(np.median(ts[2:]))

    with torch.cuda.stream(stream):
        out["restore_us"] = events(env.restore)
        out["replay_packed_device_us"] = events(lambda: env.replay_packed(d_a8, d_db, d_sb, reward=d_rw, obs=obs_ring, mask=mask_ring), pre=env.restore)
        out["replay_host_packed_kernel_us"] = events(lambda: env.replay_host_packed(h_a8, h_db, h_sb, reward=h_rw, obs=obs_ring, mask=mask_ring), pre=env.restore)
        out["replay_host_packed_flags_only_kernel_us"] = events(lambda: env.replay_host_packed(h_a8, h_db, h_sb, reward_dev=d_rw, obs=obs_ring, mask=mask_ring), pre=env.restore)
        # host wall clock of the sequence the bench's e2e leg times
        for _ in range(3):
            env.restore(); env.replay_host_packed(h_a8, h_db, h_sb, reward=h_rw, obs=obs_ring, mask=mask_ring)
        stream.synchronize()
        t0 = time.perf_counter()
        R = 20
        for _ in range(R):
            env.restore(); env.replay_host_packed(h_a8, h_db, h_sb, reward=h_rw, obs=obs_ring, mask=mask_ring)
            _ = float(h_rw[T - 1, 0]) + float(h_db[0, T - 1] & 1)
        stream.synchronize()
        out["e2e_wall_us_per_episode"] = (time.perf_counter() - t0) / R * 1e6
        t0 = time.perf_counter()
        for _ in range(R):
            env.restore()
        stream.synchronize()
        out["restore_call_wall_us"] = (time.perf_counter() - t0) / R * 1e6
        h_rw2 = [h_rw, env.host_buffer((T, B), np.float32)]
        h_db2 = [h_db, env.host_buffer((tiles, T), np.uint32)]; h_sb2 = [h_sb, env.host_buffer((tiles, T), np.uint32)]

        def submit(i):
            env.restore()
            env.replay_host_packed(h_a8, h_db2[i & 1], h_sb2[i & 1], reward=h_rw2[i & 1], obs=obs_ring, mask=mask_ring, sync=False)
            ev = torch.cuda.Event(); ev.record(stream)
            return ev

        submit(0).synchronize()
        t0 = time.perf_counter()
        pending = None
        for i in range(R):
            ev = submit(i)
            if pending is not None:
                pending[0].synchronize(); _ = float(h_rw2[pending[1] & 1][T - 1, 0])
            pending = (ev, i)
        pending[0].synchronize()
        out["e2e_pipelined_wall_us_per_episode"] = (time.perf_counter() - t0) / R * 1e6
    out["e2e_env_steps_per_s"] = B * T / (out["e2e_wall_us_per_episode"] * 1e-6)
    out["e2e_pipelined_env_steps_per_s"] = B * T / (out["e2e_pipelined_wall_us_per_episode"] * 1e-6)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
