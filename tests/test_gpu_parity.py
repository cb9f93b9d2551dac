"""GPU parity: the fused CUDA step (through the C ABI) against the CPU oracle on identical seeded targets,
action streams and injected coin / permutation draws.  Bit-exact: states, observations, masks, rewards
(f32 bit patterns), done/success flags, metric counters and solutions."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


def run_parity(name, B=192, T=40, seed=1234, add_inverts=False, add_perms=False, invalid_rate=0.05, **extra):
    from qiskit_gym_b200 import BatchedEnv

    kind, n, gateset, kw = H.config_table()[name]
    kw = dict(kw, **extra)
    rng = np.random.Generator(np.random.PCG64(seed))
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = add_inverts
    cfg = H.make_cfg(kind, n, gateset, add_perms=add_perms, **pk)
    tarr = H.random_targets(kind, n, gateset, B, seed, scramble=24, num_rotations=kw.get("max_rotations", 5) + 1, vary_rotations=True)
    lens = H.payload_lengths(kind, n, tarr)
    A = len(gateset)
    actions = H.random_actions(rng, T, B, A, invalid_rate)
    coins = rng.integers(0, 2, size=(T, B)).astype(np.uint8) if (add_inverts and kind != H.PAULI) else None
    perm_raw = rng.integers(0, 2**32, size=(T + 1, B), dtype=np.uint64).astype(np.uint32) if (kind == H.PAULI and add_perms) else None
    ref = orc.run_batch(cfg, tarr, lens, actions, coins=coins, perm_raw=perm_raw)

    env = BatchedEnv(kind, n, gateset, B, add_perms=add_perms, **pk)
    dev = env.device
    env.set_state(tarr)

    def raw(t):
        if perm_raw is None:
            return None
        return torch.from_numpy(perm_raw[t].view(np.int32).copy()).to(dev)

    obs0 = env.observe(perm_raw=raw(0)).reshape(B, -1).cpu().numpy()
    assert np.array_equal(obs0.astype(np.uint8), ref["obs0"]), f"{name}: observation after set_state differs"
    _, done0, succ0, depth0 = env.status()
    for t in range(T):
        a = torch.from_numpy(actions[t]).to(dev)
        c = None if coins is None else torch.from_numpy(coins[t]).to(dev)
        env.step(a, coins=c, perm_raw=raw(t + 1))
        obs = env.obs.reshape(B, -1).cpu().numpy()
        assert set(np.unique(obs)) <= {0.0, 1.0}
        bad = np.nonzero((obs.astype(np.uint8) != ref["obs"][t]).any(axis=1))[0]
        assert bad.size == 0, f"{name}: obs differs at step {t} for envs {bad[:8]}"
        rw = env.reward.cpu().numpy().view(np.uint32)
        assert np.array_equal(rw, ref["reward"][t].view(np.uint32)), f"{name}: reward bits differ at step {t}"
        assert np.array_equal(env.done.cpu().numpy().astype(np.uint8), ref["done"][t]), f"{name}: done differs at step {t}"
        assert np.array_equal(env.success.cpu().numpy().astype(np.uint8), ref["success"][t]), f"{name}: success differs at step {t}"
        mask = env.mask.cpu().numpy().astype(np.uint8)
        assert np.array_equal(mask, np.repeat((1 - ref["success"][t])[:, None], A, axis=1)), f"{name}: mask differs at step {t}"
        if t % 7 == 0 or t == T - 1:
            m = env.metrics().cpu().numpy().astype(np.int64)
            assert np.array_equal(m, ref["counts"][t]), f"{name}: metric counters differ at step {t}"
            _, _, _, depth = env.status()
            assert np.array_equal(depth.cpu().numpy().astype(np.int64), ref["depth"][t])
    assert int(env.errors().max().item()) & ~16 == 0
    for b in range(0, B, max(1, B // 48)):
        st = env.get_state(b)
        L = int(ref["final_state_len"][b])
        assert np.array_equal(st, ref["final_state"][b, :L]), f"{name}: final state differs for env {b}"
        sol = env.solution(b)
        SL = int(ref["sol_len"][b])
        assert sol == ref["solutions"][b, :SL].tolist(), f"{name}: solution differs for env {b}"


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C5_perm27_heavyhex", "lf5_line_swap",
                                  "lf11_line", "clifford3_allgates", "clifford5_allgates", "perm5_mixed"])
def test_step_parity(name):
    run_parity(name)


@pytest.mark.parametrize("tile", [16, 32])
@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C4_pauli10_line", "C5_perm27_heavyhex", "pauli6_line",
                                  "clifford20_line", "lf11_line", "perm5_mixed"])
def test_parity_for_both_tile_sizes(name, tile):
    """qg_config.tile_envs: 16- and 32-env warp tiles (what a launch picks from the batch size when it is 0) give the same bits, single
    steps and replay, with and without the add_inverts inverse; B = 77 leaves a ragged last tile for both."""
    kind = H.config_table()[name][0]
    run_parity(name, B=77, T=20, seed=31, tile_envs=tile, add_inverts=(kind != H.PAULI), add_perms=(kind == H.PAULI),
               invalid_rate=0.0 if kind == H.PAULI else 0.05)          # (the reference panics on an out-of-range action under add_perms, pauli.rs:594-599)
    run_replay_parity(name, B=77, T=35, seed=32, tile_envs=tile, ring=3)


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C5_perm27_heavyhex", "lf5_line_swap",
                                  "clifford5_allgates", "perm5_mixed"])
def test_step_parity_with_inverts(name):
    run_parity(name, add_inverts=True, seed=77)


def run_replay_parity(name, B=200, T=40, seed=4321, add_inverts=False, add_perms=False, invalid_rate=0.05, ring=None, **extra):
    """qg_replay (T fused steps in one launch) must deliver, step by step, exactly what the oracle's T single steps do."""
    from qiskit_gym_b200 import BatchedEnv

    kind, n, gateset, kw = H.config_table()[name]
    kw = dict(kw, **extra)
    rng = np.random.Generator(np.random.PCG64(seed))
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = add_inverts
    cfg = H.make_cfg(kind, n, gateset, add_perms=add_perms, **pk)
    tarr = H.random_targets(kind, n, gateset, B, seed, scramble=24, num_rotations=kw.get("max_rotations", 5) + 1, vary_rotations=True)
    lens = H.payload_lengths(kind, n, tarr)
    A = len(gateset)
    actions = H.random_actions(rng, T, B, A, invalid_rate)
    coins = rng.integers(0, 2, size=(T, B)).astype(np.uint8) if (add_inverts and kind != H.PAULI) else None
    perm_raw = rng.integers(0, 2**32, size=(T + 1, B), dtype=np.uint64).astype(np.uint32) if (kind == H.PAULI and add_perms) else None
    ref = orc.run_batch(cfg, tarr, lens, actions, coins=coins, perm_raw=perm_raw)

    env = BatchedEnv(kind, n, gateset, B, add_perms=add_perms, **pk)
    dev = env.device
    env.set_state(tarr)
    if perm_raw is not None:
        env.observe(perm_raw=torch.from_numpy(perm_raw[0].view(np.int32).copy()).to(dev))     # observe() after set_state picks perm 0
    R = T if ring is None else ring
    osz = int(np.prod(env.obs_shape()))
    obs = torch.full((R, B, osz), 7.0, dtype=torch.float32, device=dev)
    mask = torch.zeros((R, B, A), dtype=torch.bool, device=dev)
    reward = torch.zeros((T, B), dtype=torch.float32, device=dev)
    done = torch.zeros((T, B), dtype=torch.bool, device=dev)
    success = torch.zeros((T, B), dtype=torch.bool, device=dev)
    env.replay(torch.from_numpy(actions).to(dev), coins=None if coins is None else torch.from_numpy(coins).to(dev),
               perm_raw=None if perm_raw is None else torch.from_numpy(perm_raw[1:].view(np.int32).copy()).to(dev),
               obs=obs, mask=mask, reward=reward, done=done, success=success)
    assert np.array_equal(reward.cpu().numpy().view(np.uint32), ref["reward"].view(np.uint32)), f"{name}: replay rewards differ"
    assert np.array_equal(done.cpu().numpy().astype(np.uint8), ref["done"]) and np.array_equal(success.cpu().numpy().astype(np.uint8), ref["success"])
    ob = obs.cpu().numpy()
    mk = mask.cpu().numpy().astype(np.uint8)
    for t in range(max(0, T - R), T):
        assert np.array_equal(ob[t % R].astype(np.uint8), ref["obs"][t]), f"{name}: replay obs differs at step {t}"
        assert np.array_equal(mk[t % R], np.repeat((1 - ref["success"][t])[:, None], A, axis=1)), f"{name}: replay mask differs at step {t}"
    assert np.array_equal(env.metrics().cpu().numpy().astype(np.int64), ref["counts"][T - 1])
    _, _, _, depth = env.status()
    assert np.array_equal(depth.cpu().numpy().astype(np.int64), ref["depth"][T - 1])
    for b in range(0, B, max(1, B // 40)):
        L = int(ref["final_state_len"][b])
        assert np.array_equal(env.get_state(b), ref["final_state"][b, :L]), f"{name}: replay final state differs for env {b}"
        assert env.solution(b) == ref["solutions"][b, : int(ref["sol_len"][b])].tolist(), f"{name}: replay solution differs for env {b}"
    # a replayed batch continues exactly like a stepped one
    env.step(torch.zeros(B, dtype=torch.int32, device=dev), coins=None if coins is None else torch.zeros(B, dtype=torch.uint8, device=dev),
             perm_raw=None if perm_raw is None else torch.zeros(B, dtype=torch.int32, device=dev))


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C4_pauli10_line", "C5_perm27_heavyhex", "lf5_line_swap",
                                  "lf11_line", "clifford3_allgates", "pauli3_line", "perm5_mixed"])
def test_replay_parity(name):
    run_replay_parity(name)


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C3_clifford8_full", "lf5_line_swap", "clifford20_line"])
def test_replay_parity_with_inverts(name):
    run_replay_parity(name, add_inverts=True, T=24, seed=11)


def test_replay_ring_and_perms():
    run_replay_parity("C3_clifford8_full", B=97, T=20, ring=3)
    run_replay_parity("C1_perm_grid3", B=33, T=9, ring=2)
    run_replay_parity("pauli6_line", B=70, T=30, add_perms=True, invalid_rate=0.0, seed=8)
    run_replay_parity("C4_pauli10_line", B=40, T=12, add_perms=True, invalid_rate=0.0, ring=5, seed=9)


@pytest.mark.parametrize("name,B,T,inv", [("C3_clifford8_full", 5000, 70, False), ("C1_perm_grid3", 300, 40, True), ("C4_pauli10_line", 257, 33, False)])
def test_replay_host_pipeline(name, B, T, inv):
    """qg_replay_host: chunked, pipelined episode replay with host buffers == the oracle's step-by-step results."""
    from qiskit_gym_b200 import BatchedEnv

    kind, n, gateset, kw = H.config_table()[name]
    rng = np.random.Generator(np.random.PCG64(21))
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = inv
    cfg = H.make_cfg(kind, n, gateset, add_perms=False, **pk)
    tarr = H.random_targets(kind, n, gateset, B, 21, scramble=30)
    lens = H.payload_lengths(kind, n, tarr)
    A = len(gateset)
    actions = H.random_actions(rng, T, B, A, 0.02)
    coins = rng.integers(0, 2, size=(T, B)).astype(np.uint8) if inv else None
    ref = orc.run_batch(cfg, tarr, lens, actions, coins=coins, want_obs=True)
    env = BatchedEnv(kind, n, gateset, B, add_perms=False, **pk)
    env.set_state(tarr)
    osz = int(np.prod(env.obs_shape()))
    ring = 3
    obs = torch.zeros((ring, B, osz), dtype=torch.float32, device=env.device)
    mask = torch.zeros((ring, B, A), dtype=torch.bool, device=env.device)
    a_pin = torch.from_numpy(actions).pin_memory()
    rew = torch.zeros((T, B), dtype=torch.float32).pin_memory(); don = torch.zeros((T, B), dtype=torch.uint8).pin_memory(); suc = torch.zeros((T, B), dtype=torch.uint8).pin_memory()
    env.replay_host(a_pin.numpy(), rew.numpy(), don.numpy(), suc.numpy(), coins=coins, obs=obs, mask=mask)
    assert np.array_equal(rew.numpy().view(np.uint32), ref["reward"].view(np.uint32))
    assert np.array_equal(don.numpy(), ref["done"]) and np.array_equal(suc.numpy(), ref["success"])
    ob = obs.cpu().numpy().astype(np.uint8)
    for t in range(T - ring, T):
        assert np.array_equal(ob[t % ring], ref["obs"][t]), f"{name}: ring slot of step {t} differs"
    for b in range(0, B, max(1, B // 25)):
        assert env.solution(b) == ref["solutions"][b, : int(ref["sol_len"][b])].tolist()
    # pageable host memory works too (copies are staged by the runtime)
    env.set_state(tarr)
    r2 = np.zeros((T, B), dtype=np.float32)
    env.replay_host(actions, r2, coins=coins)
    assert np.array_equal(r2.view(np.uint32), ref["reward"].view(np.uint32))


@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_pinned_and_pageable(pinned):
    """qg_step_host: with pinned buffers the kernel reads / writes host memory itself (zero copy), with pageable ones the call stages
    copies; both must deliver what the oracle's steps do."""
    from qiskit_gym_b200 import BatchedEnv

    kind, n, gateset, kw = H.config_table()["C3_clifford8_full"]
    B, T = 777, 12
    rng = np.random.Generator(np.random.PCG64(31))
    cfg = H.make_cfg(kind, n, gateset, add_inverts=True, add_perms=False)
    tarr = H.random_targets(kind, n, gateset, B, 31, scramble=20)
    actions = H.random_actions(rng, T, B, len(gateset), 0.02)
    coins = rng.integers(0, 2, size=(T, B)).astype(np.uint8)
    ref = orc.run_batch(cfg, tarr, H.payload_lengths(kind, n, tarr), actions, coins=coins)
    env = BatchedEnv(kind, n, gateset, B, add_inverts=True, add_perms=False)
    env.set_state(tarr)

    def buf(shape, dt):
        t = torch.zeros(shape, dtype=dt)
        return (t.pin_memory() if pinned else t).numpy()
    a_h, c_h = buf((T, B), torch.int32), buf((T, B), torch.uint8)
    a_h[:] = actions; c_h[:] = coins
    rew, don, suc = buf((B,), torch.float32), buf((B,), torch.uint8), buf((B,), torch.uint8)
    for t in range(T):
        env.step_host(a_h[t], rew, don, suc, coins=c_h[t], obs=env.obs)
        assert np.array_equal(rew.view(np.uint32), ref["reward"][t].view(np.uint32)), f"reward bits differ at step {t}"
        assert np.array_equal(don, ref["done"][t]) and np.array_equal(suc, ref["success"][t])
        assert np.array_equal(env.obs.reshape(B, -1).cpu().numpy().astype(np.uint8), ref["obs"][t])


@pytest.mark.parametrize("name", ["lf4_swap", "lf8_swap", "lf32_line", "clifford4_allgates", "clifford8_allgates", "clifford16_allgates"])
@pytest.mark.parametrize("inv", [False, True])
def test_step_parity_power_of_two_rows(name, inv):
    """Row widths 4 .. 32 (a row inside one word: row_xor_pow2 / row_swap_pow2) with every gate kind, single steps and replay."""
    run_parity(name, B=130, T=30, add_inverts=inv, seed=9)
    run_replay_parity(name, B=70, T=24, add_inverts=inv, seed=10)


@pytest.mark.parametrize("name", ["clifford20_line", "lf40_line"])
def test_step_parity_wide_rows(name):
    run_parity(name, B=70, T=24, add_inverts=True, seed=5)


@pytest.mark.parametrize("name", ["C4_pauli10_line", "pauli3_line", "pauli6_line"])
def test_pauli_parity(name):
    run_parity(name, T=60)


@pytest.mark.parametrize("name", ["C4_pauli10_line", "pauli3_line", "pauli6_line"])
def test_pauli_parity_with_perms(name):
    # no out-of-range actions here: the reference indexes act_perms[perm][action] and panics (pauli.rs:594-599)
    run_parity(name, T=40, add_perms=True, seed=99, invalid_rate=0.0)


def test_ragged_batch_sizes():
    for B in (1, 31, 33, 65, 100):
        run_parity("C3_clifford8_full", B=B, T=6, seed=B)
        run_parity("C1_perm_grid3", B=B, T=6, seed=B)
        run_parity("C2_lf8_line", B=B, T=6, seed=B)
        run_parity("lf5_line_swap", B=B, T=6, seed=B)       # 25 observation entries, 16 actions
        run_parity("clifford3_allgates", B=B, T=6, seed=B)


def test_large_permutation_direct_expander():
    """n > 64: no observation bit stream in shared memory, the expander tests the packed bytes (one-hot rows)."""
    from qiskit_gym_b200 import BatchedEnv
    n, B, T = 70, 37, 12
    gs = [("SWAP", (i, i + 1)) for i in range(n - 1)]
    rng = np.random.Generator(np.random.PCG64(3))
    tarr = np.stack([rng.permutation(n) for _ in range(B)]).astype(np.int64)
    actions = H.random_actions(rng, T, B, len(gs), 0.05)
    cfg = H.make_cfg(H.PERM, n, gs, add_inverts=False, add_perms=False)
    ref = orc.run_batch(cfg, tarr, H.payload_lengths(H.PERM, n, tarr), actions)
    env = BatchedEnv(H.PERM, n, gs, B, add_inverts=False, add_perms=False)
    env.set_state(tarr)
    for t in range(T):
        env.step(torch.from_numpy(actions[t]).to(env.device))
        assert np.array_equal(env.obs.reshape(B, -1).cpu().numpy().astype(np.uint8), ref["obs"][t])
        assert np.array_equal(env.reward.cpu().numpy().view(np.uint32), ref["reward"][t].view(np.uint32))


@pytest.mark.parametrize("name,difficulty", [("C1_perm_grid3", 7), ("C2_lf8_line", 40), ("C3_clifford8_full", 256), ("C5_perm27_heavyhex", 100),
                                              ("clifford5_allgates", 33), ("lf40_line", 64), ("C4_pauli10_line", 48), ("pauli6_line", 80),
                                              ("pauli3_line", 20), ("C4_pauli10_line", 0)])
def test_reset_parity(name, difficulty):
    """Env::reset with the engine's Philox stream injected into the oracle: states, depth, success, observation."""
    from qiskit_gym_b200 import BatchedEnv

    kind, n, gateset, kw = H.config_table()[name]
    B, seed, first = 300, 0xC0FFEE + difficulty, 1000
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = False
    env = BatchedEnv(kind, n, gateset, B, difficulty=difficulty, add_perms=False, **pk)
    env.reset(seed=seed, first_env_id=first)
    obs = env.observe().reshape(B, -1).cpu().numpy().astype(np.uint8)
    reward, done, success, depth = [t.cpu().numpy() for t in env.status()]
    ref = orc.OracleEnv(kind, n, gateset, difficulty=difficulty, add_perms=False, **pk)
    assert int(env.errors().max().item()) == 0
    for b in range(0, B, 3):
        ref.reset(seed=seed, env_id=first + b)
        assert np.array_equal(env.get_state(b), ref.raw_state()), f"{name}: reset state differs for env {b}"
        d = np.zeros(obs.shape[1], dtype=np.uint8)
        d[ref.observe()] = 1
        assert np.array_equal(obs[b], d), f"{name}: reset observation differs for env {b}"
        assert int(depth[b]) == ref.depth() and bool(success[b]) == ref.success() and bool(done[b]) == ref.is_final()
        assert float(reward[b]) == ref.reward()
    # a reset batch must step exactly like the oracle afterwards (metrics / DAG order were re-initialised)
    rng = np.random.Generator(np.random.PCG64(1))
    acts = H.random_actions(rng, 10, B, len(gateset))
    refs = []
    for b in range(0, B, 25):
        r = orc.OracleEnv(kind, n, gateset, difficulty=difficulty, add_perms=False, **pk)
        r.reset(seed=seed, env_id=first + b)
        refs.append((b, r))
    for t in range(10):
        env.step(torch.from_numpy(acts[t]).to(env.device))
        rw = env.reward.cpu().numpy()
        for b, r in refs:
            r.step(int(acts[t, b]))
            assert pyref_bits(rw[b]) == pyref_bits(r.reward())
    for b, r in refs:
        assert np.array_equal(env.get_state(b), r.raw_state()) and env.solution(b) == r.solution()


def pyref_bits(x):
    import struct
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def test_single_env_classes_match_notebook():
    """The drop-in raw-env classes (batch of one) reproduce the reference notebook's LinearFunction walk-through."""
    import json
    import os
    from qiskit_gym_b200 import LinearFunctionEnv, PermutationEnv

    nb = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_kats.json")))
    gs = [(g, tuple(q)) for g, q in nb["lf3_gateset"]]
    env = LinearFunctionEnv(3, 1, gs, 2, 128, add_inverts=False)
    assert env.num_actions() == nb["lf3_action_space"] and env.obs_shape() == nb["lf3_obs_space"]
    assert env.is_final() and env.reward() == 1.0 and env.observe() == [0, 4, 8]      # constructor state: identity, success
    for key in ("lf3_sequence_a", "lf3_sequence_b"):
        seq = nb[key]
        env.set_state(np.array(seq["start"]).reshape(-1).tolist())
        for a, st, fin in zip(seq["actions"], seq["states"], seq["is_final"]):
            env.step(a)
            d = np.zeros(9, dtype=int); d[env.observe()] = 1
            assert d.reshape(3, 3).tolist() == st and env.is_final() == fin
        assert env.solution() == seq["actions"] and env.masks() == [False] * 8
    obs_perms, act_perms = env.twists()
    assert len(obs_perms) == 2 and len(act_perms[0]) == 8
    env.difficulty = 5
    assert env.difficulty == 5
    env.reset(seed=3)
    assert not env.is_final() or env.success()
    p = PermutationEnv(9, 1, [(g, tuple(q)) for g, q in nb["perm_grid3_gateset"]], 2, 128)
    assert p.obs_shape() == [9, 9] and p.num_actions() == 12 and len(p.twists()[0]) == 8
