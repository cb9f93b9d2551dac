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


def _device_digest(obs_all, mask_all, rew, done, succ, wobs, wmask):
    """The oracle's per-env digest (oracle/qg_oracle_c.cpp: qgo_digest) formed from the engine's output tensors: int64 arithmetic wraps like uint64."""
    dev = obs_all.device
    T, B = rew.shape
    wo = torch.from_numpy(wobs.astype(np.int64)).to(dev)
    wm = torch.from_numpy(wmask.astype(np.int64)).to(dev)
    d = torch.zeros(B, dtype=torch.int64, device=dev)
    for t in range(T):
        h = (obs_all[t].to(torch.int64) * wo).sum(1) + (mask_all[t].to(torch.int64) * wm).sum(1)
        rb = rew[t].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        h = h + rb * 0x9E3779B1 + done[t].to(torch.int64) * 0x85EBCA6B + succ[t].to(torch.int64) * 0xC2B2AE35
        d = d * 0x100000001B3 + h
    return d


@pytest.mark.parametrize("name,inverts", [("C1_perm_grid3", False), ("C2_lf8_line", False), ("C3_clifford8_full", False), ("C4_pauli10_line", False),
                                          ("C5_perm27_heavyhex", False), ("C2_lf8_line", True), ("C3_clifford8_full", True), ("C1_perm_grid3", True)])
def test_whole_batch_digest_matches_oracle_full_size(name, inverts):
    """EVERY environment of the full 65 536 x 128 batch against the oracle: every observation entry, mask entry, reward bit pattern and flag of
    every step, folded into one uint64 per environment on both sides (the oracle runs on all host threads; one observation slab per step is
    kept on the device, as in bench.py)."""
    import os
    env, kind, n, gateset, kw = _env(name, B_FULL, **({"add_inverts": True} if inverts else {}))
    dev, A = env.device, len(gateset)
    tarr = _targets(kind, n, gateset, B_FULL, 17, kw)
    rng = np.random.Generator(np.random.PCG64(123))
    actions = H.random_actions(rng, T_FULL, B_FULL, A, 0.01)
    osz = int(np.prod(env.obs_shape()))
    wobs = rng.integers(1, 2 ** 31, size=osz, dtype=np.uint64)
    wmask = rng.integers(1, 2 ** 31, size=A, dtype=np.uint64)
    env.set_state(tarr)
    env.observe()
    obs_all = torch.empty((T_FULL, B_FULL, osz), dtype=torch.float32, device=dev)
    mask_all = torch.empty((T_FULL, B_FULL, A), dtype=torch.bool, device=dev)
    rew = torch.empty((T_FULL, B_FULL), dtype=torch.float32, device=dev)
    done = torch.empty((T_FULL, B_FULL), dtype=torch.bool, device=dev)
    succ = torch.empty((T_FULL, B_FULL), dtype=torch.bool, device=dev)
    coins = rng.integers(0, 2, size=(T_FULL, B_FULL)).astype(np.uint8) if inverts else None       # the invert coin of every step, injected on both sides
    env.replay(torch.from_numpy(actions).to(dev), coins=None if coins is None else torch.from_numpy(coins).to(dev), obs=obs_all, mask=mask_all, reward=rew, done=done,
               success=succ)
    got = _device_digest(obs_all, mask_all, rew, done, succ, wobs, wmask).cpu().numpy().view(np.uint64)
    del obs_all, mask_all
    cfg = H.make_cfg(kind, n, gateset, add_perms=False, **kw)
    ref = orc.digest(cfg, tarr, H.payload_lengths(kind, n, tarr), actions, wobs, wmask, coins=coins, threads=max(1, min(32, os.cpu_count() or 1)))
    bad = np.nonzero(got != ref)[0]
    assert bad.size == 0, f"{name}: {bad.size} of {B_FULL} environments differ from the oracle, first {bad[:8].tolist()}"
    assert int(env.errors().max().item()) == 0


def _digest_case(name, seed):
    import os
    env, kind, n, gateset, kw = _env(name, B_FULL)
    A = len(gateset)
    tarr = _targets(kind, n, gateset, B_FULL, seed, kw)
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    actions = H.random_actions(rng, T_FULL, B_FULL, A, 0.01)
    osz = int(np.prod(env.obs_shape()))
    wobs = rng.integers(1, 2 ** 31, size=osz, dtype=np.uint64)
    wmask = rng.integers(1, 2 ** 31, size=A, dtype=np.uint64)
    cfg = H.make_cfg(kind, n, gateset, add_perms=False, **kw)
    ref = orc.digest(cfg, tarr, H.payload_lengths(kind, n, tarr), actions, wobs, wmask, threads=max(1, min(32, os.cpu_count() or 1)))
    return env, tarr, actions, osz, A, wobs, wmask, ref


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C4_pauli10_line", "C5_perm27_heavyhex"])
def test_whole_batch_digest_single_step_launches_full_size(name):
    """The same whole-batch comparison through 128 single-step launches (qg_step: the policy-in-the-loop granularity; 16-env warp tiles for the
    large observations of C4 / C5)."""
    env, tarr, actions, osz, A, wobs, wmask, ref = _digest_case(name, 19)
    dev = env.device
    env.set_state(tarr)
    env.observe()
    a_dev = torch.from_numpy(actions).to(dev)
    obs_t = torch.empty((B_FULL, osz), dtype=torch.float32, device=dev)
    mask_t = torch.empty((B_FULL, A), dtype=torch.bool, device=dev)
    wo = torch.from_numpy(wobs.astype(np.int64)).to(dev)
    wm = torch.from_numpy(wmask.astype(np.int64)).to(dev)
    d = torch.zeros(B_FULL, dtype=torch.int64, device=dev)
    for t in range(T_FULL):
        env.step(a_dev[t], obs=obs_t, mask=mask_t)
        d = _device_digest(obs_t[None], mask_t[None], env.reward[None], env.done[None], env.success[None], wobs, wmask) + d * 0x100000001B3
    got = d.cpu().numpy().view(np.uint64)
    bad = np.nonzero(got != ref)[0]
    assert bad.size == 0, f"{name}: {bad.size} of {B_FULL} environments differ from the oracle, first {bad[:8].tolist()}"


@pytest.mark.parametrize("name", ["C3_clifford8_full", "C4_pauli10_line", "C1_perm_grid3"])
def test_whole_batch_digest_packed_observations_full_size(name):
    """... and through the packed-bit observation path (qg_replay_bits: what the fused policy kernels read), bits unpacked on the device."""
    env, tarr, actions, osz, A, wobs, wmask, ref = _digest_case(name, 23)
    dev = env.device
    env.set_state(tarr)
    env.observe()
    W = env.obs_words()
    bits = torch.empty((T_FULL, B_FULL, W), dtype=torch.int32, device=dev)
    mask_all = torch.empty((T_FULL, B_FULL, A), dtype=torch.bool, device=dev)
    rew = torch.empty((T_FULL, B_FULL), dtype=torch.float32, device=dev)
    done = torch.empty((T_FULL, B_FULL), dtype=torch.bool, device=dev)
    succ = torch.empty((T_FULL, B_FULL), dtype=torch.bool, device=dev)
    env.replay_bits(torch.from_numpy(actions).to(dev), obs_bits=bits, mask=mask_all, reward=rew, done=done, success=succ)
    sh = torch.arange(32, dtype=torch.int32, device=dev)
    d = torch.zeros(B_FULL, dtype=torch.int64, device=dev)
    for t in range(T_FULL):
        dense = ((bits[t][:, :, None] >> sh) & 1).reshape(B_FULL, W * 32)[:, :osz].to(torch.float32)
        d = _device_digest(dense[None], mask_all[t][None], rew[t][None], done[t][None], succ[t][None], wobs, wmask) + d * 0x100000001B3
    got = d.cpu().numpy().view(np.uint64)
    bad = np.nonzero(got != ref)[0]
    assert bad.size == 0, f"{name}: {bad.size} of {B_FULL} environments differ from the oracle, first {bad[:8].tolist()}"
