"""GPU parity at BASELINE.json's full sizes (65 536 environments x 128 env-steps per config), where the oracle cannot run the whole
batch in seconds: size-independent properties plus an oracle check of a scattered sample.

  * involution round trip — on the GF(2) state every gate of every gateset is its own inverse (a row XOR twice, a row swap twice; S
    and SX act on the phase-free tableau as row XORs: clifford.rs:89-133, linear_function.rs:62-83, permutation.rs:205-208), so
    playing an action stream forwards and then backwards must bring every one of the 65 536 environments back to its target;
  * replay == steps — one launch that plays 128 steps equals 128 single-step launches, output for output;
  * independence — an environment's results do not depend on the batch it sits in (what sharding over GPUs relies on): a scattered
    sample re-run as its own small batch, and run by the CPU oracle, gives the same rewards / flags / observations / counters.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

B_FULL, T_FULL = 65536, 128


def _targets(kind, n, gateset, B, seed, kw):
    base = min(B, 8192)                     # host-side scrambling of 65 536 matrices is slow; tile a seeded set, shifted per copy
    t = H.random_targets(kind, n, gateset, base, seed, scramble=64, num_rotations=kw.get("max_rotations", 5))
    idx = (np.arange(B) * 2654435761 % base).astype(np.int64)
    return t[idx]


def _env(name, B, **extra):
    from qiskit_gym_b200 import BatchedEnv
    kind, n, gateset, kw = H.config_table()[name]
    pk = dict(kw, **extra)
    if kind != H.PAULI:
        pk.setdefault("add_inverts", False)
    return BatchedEnv(kind, n, gateset, B, add_perms=False, **pk), kind, n, gateset, pk


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C5_perm27_heavyhex"])
def test_forward_backward_round_trip_full_size(name):
    env, kind, n, gateset, kw = _env(name, B_FULL)
    dev, A = env.device, len(gateset)
    tarr = _targets(kind, n, gateset, B_FULL, 11, kw)
    env.set_state(tarr)
    obs0 = env.observe().clone()
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    actions = torch.randint(0, A, (T_FULL, B_FULL), dtype=torch.int32, device=dev, generator=g)
    reward = torch.empty((T_FULL, B_FULL), dtype=torch.float32, device=dev)
    env.replay(actions, reward=reward)
    mid = env.observe().clone()
    assert not torch.equal(mid, obs0)                                     # the stream did move the batch
    env.replay(torch.flip(actions, dims=[0]).contiguous(), reward=reward)
    back = env.observe()
    assert torch.equal(back, obs0), f"{name}: {int((back != obs0).flatten(1).any(1).sum())} of {B_FULL} environments did not return to their target"
    # 2 x 128 gates were counted for every environment (metrics.rs:64-123; SWAP counts 3 CX, permutation.rs logs every valid action)
    m = env.metrics().cpu().numpy().astype(np.int64)
    per_gate = 3 if gateset[0][0] == "SWAP" else 1
    if all(gname == gateset[0][0] for gname, _ in gateset):
        assert np.all(m[:, 3] == 2 * T_FULL * per_gate)
    assert int(env.errors().max().item()) & ~2 == 0            # 2 = solution log full: 256 steps were played into a log of max_depth entries


@pytest.mark.parametrize("name", ["C3_clifford8_full", "C4_pauli10_line"])
def test_replay_equals_steps_full_size(name):
    env, kind, n, gateset, kw = _env(name, B_FULL)
    dev, A = env.device, len(gateset)
    T = 32
    tarr = _targets(kind, n, gateset, B_FULL, 12, kw)
    g = torch.Generator(device=dev)
    g.manual_seed(6)
    actions = torch.randint(0, A, (T, B_FULL), dtype=torch.int32, device=dev, generator=g)
    osz = int(np.prod(env.obs_shape()))
    env.set_state(tarr)
    env.observe()
    obs_r = torch.empty((2, B_FULL, osz), dtype=torch.float32, device=dev)
    mask_r = torch.empty((2, B_FULL, A), dtype=torch.bool, device=dev)
    rew_r = torch.empty((T, B_FULL), dtype=torch.float32, device=dev)
    done_r = torch.empty((T, B_FULL), dtype=torch.bool, device=dev)
    env.replay(actions, obs=obs_r, mask=mask_r, reward=rew_r, done=done_r)
    met_r = env.metrics().clone()
    env.set_state(tarr)
    env.observe()
    for t in range(T):
        env.step(actions[t])
        assert torch.equal(env.reward.view(torch.int32), rew_r[t].view(torch.int32)), f"{name}: reward bits differ at step {t}"
        assert torch.equal(env.done, done_r[t])
    assert torch.equal(env.obs.reshape(B_FULL, -1), obs_r[(T - 1) % 2]) and torch.equal(env.mask, mask_r[(T - 1) % 2])
    assert torch.equal(env.metrics(), met_r)


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C4_pauli10_line", "C5_perm27_heavyhex"])
def test_scattered_sample_matches_oracle_and_small_batch_full_size(name):
    env, kind, n, gateset, kw = _env(name, B_FULL)
    dev, A = env.device, len(gateset)
    tarr = _targets(kind, n, gateset, B_FULL, 13, kw)
    rng = np.random.Generator(np.random.PCG64(99))
    actions = H.random_actions(rng, T_FULL, B_FULL, A, 0.01)
    osz = int(np.prod(env.obs_shape()))
    env.set_state(tarr)
    env.observe()
    rew = torch.empty((T_FULL, B_FULL), dtype=torch.float32, device=dev)
    done = torch.empty((T_FULL, B_FULL), dtype=torch.bool, device=dev)
    succ = torch.empty((T_FULL, B_FULL), dtype=torch.bool, device=dev)
    obs = torch.empty((1, B_FULL, osz), dtype=torch.float32, device=dev)
    env.replay(torch.from_numpy(actions).to(dev), obs=obs, reward=rew, done=done, success=succ)
    # the sample: first / last environments, tile edges and random picks
    pick = np.unique(np.concatenate([[0, 1, 31, 32, 33, 63, 64, B_FULL - 33, B_FULL - 32, B_FULL - 1], rng.integers(0, B_FULL, size=182)])).astype(np.int64)
    S = pick.size
    cfg = H.make_cfg(kind, n, gateset, add_perms=False, **kw)
    sub_t, sub_a = np.ascontiguousarray(tarr[pick]), np.ascontiguousarray(actions[:, pick])
    ref = orc.run_batch(cfg, sub_t, H.payload_lengths(kind, n, sub_t), sub_a)
    pk = torch.from_numpy(pick).to(dev)
    assert np.array_equal(rew[:, pk].cpu().numpy().view(np.uint32), ref["reward"].view(np.uint32)), f"{name}: reward bits differ from the oracle"
    assert np.array_equal(done[:, pk].cpu().numpy().astype(np.uint8), ref["done"]) and np.array_equal(succ[:, pk].cpu().numpy().astype(np.uint8), ref["success"])
    assert np.array_equal(obs[0, pk].cpu().numpy().astype(np.uint8), ref["obs"][T_FULL - 1]), f"{name}: final observations differ from the oracle"
    assert np.array_equal(env.metrics()[pk].cpu().numpy().astype(np.int64), ref["counts"][T_FULL - 1])
    for j in range(0, S, 16):
        b = int(pick[j])
        assert env.solution(b) == ref["solutions"][j, : int(ref["sol_len"][j])].tolist(), f"{name}: solution of env {b} differs from the oracle"
    # the same environments as their own small batch (a different tiling of warps / CTAs, as on another GPU's shard)
    small, *_ = _env(name, S)
    small.set_state(sub_t)
    small.observe()
    rew_s = torch.empty((T_FULL, S), dtype=torch.float32, device=dev)
    obs_s = torch.empty((1, S, osz), dtype=torch.float32, device=dev)
    small.replay(torch.from_numpy(sub_a).to(dev), obs=obs_s, reward=rew_s)
    assert torch.equal(rew_s.view(torch.int32), rew[:, pk].view(torch.int32)) and torch.equal(obs_s[0], obs[0, pk])
