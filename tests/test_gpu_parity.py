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


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C5_perm27_heavyhex", "lf5_line_swap",
                                  "clifford5_allgates", "perm5_mixed"])
def test_step_parity_with_inverts(name):
    run_parity(name, add_inverts=True, seed=77)


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
