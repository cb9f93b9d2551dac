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
