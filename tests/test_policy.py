"""Packed-bit observations (qg_*_bits) against the dense observation and the oracle, and the fused action network
(qg_policy_forward_bits) against the PyTorch BasicPolicy it was built from."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H


def unpack_bits(bits, n):
    """int32/uint32 [B, words] -> uint8 [B, n]"""
    b = np.ascontiguousarray(bits).view(np.uint32)
    out = np.unpackbits(b.view(np.uint8).reshape(b.shape[0], -1), axis=1, bitorder="little")
    return out[:, :n]


def test_pack_obs_bits_helper():
    from qiskit_gym_b200.policy import pack_obs_bits
    rng = np.random.Generator(np.random.PCG64(1))
    for B, n in ((1, 1), (3, 31), (5, 32), (4, 33), (7, 81), (2, 729)):
        dense = (rng.random((B, n)) < 0.4).astype(np.float32)
        packed = pack_obs_bits(torch.from_numpy(dense)).numpy()
        assert packed.shape == (B, (n + 31) // 32)
        assert np.array_equal(unpack_bits(packed, n), dense.astype(np.uint8))
        # bits beyond n are zero
        assert np.array_equal(unpack_bits(packed, packed.shape[1] * 32)[:, n:], np.zeros((B, packed.shape[1] * 32 - n), np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C4_pauli10_line", "C5_perm27_heavyhex", "lf11_line", "clifford5_allgates", "pauli6_line"])
@pytest.mark.parametrize("B", [1, 33, 100])
def test_packed_observation_equals_dense_and_oracle(name, B):
    from qiskit_gym_b200 import BatchedEnv
    kind, n, gs, kw = H.config_table()[name]
    kw = dict(kw, add_perms=False)
    if kind != H.PAULI:
        kw["add_inverts"] = False
    T = 6
    rng = np.random.Generator(np.random.PCG64(17))
    cfg = H.make_cfg(kind, n, gs, **kw)
    tarr = H.random_targets(kind, n, gs, B, 5, scramble=12)
    lens = H.payload_lengths(kind, n, tarr)
    actions = H.random_actions(rng, T, B, len(gs))
    ref = orc.run_batch(cfg, tarr, lens, actions)
    env = BatchedEnv(kind, n, gs, B, **kw)
    env.set_state(tarr)
    size = env._obs_size
    bits = env.new_obs_bits()
    env.observe_bits(bits)
    assert np.array_equal(unpack_bits(bits.cpu().numpy(), size), ref["obs0"])
    for t in range(T):
        env.step_bits(torch.from_numpy(actions[t]).to(env.device), bits, mask=True)
        got = bits.cpu().numpy()
        assert np.array_equal(unpack_bits(got, size), ref["obs"][t]), (name, t)
        assert np.array_equal(unpack_bits(got, got.shape[1] * 32)[:, size:].sum(), 0)          # padding bits stay clear
        assert np.array_equal(env.reward.cpu().numpy().view(np.uint32), ref["reward"][t].view(np.uint32))
        assert np.array_equal(env.done.cpu().numpy().astype(np.uint8), ref["done"][t])
    # the same episode in one launch, packed ring of 2 slots
    env.set_state(tarr)
    ring = env.new_obs_bits(ring=2)
    env.replay_bits(torch.from_numpy(actions).to(env.device), obs_bits=ring)
    got = ring.cpu().numpy()
    assert np.array_equal(unpack_bits(got[(T - 1) % 2], size), ref["obs"][T - 1])
    assert np.array_equal(unpack_bits(got[(T - 2) % 2], size), ref["obs"][T - 2])


@pytest.mark.gpu
@pytest.mark.parametrize("obs_shape,A,emb,common,pol_layers,density", [
    ((9, 9), 12, 64, (32,), (), 0.12),
    ((27, 27), 28, 512, (256,), (), 0.04),
    ((16, 16), 72, 512, (256,), (), 0.5),
    ((20, 25), 104, 300, (200, 100), (50,), 0.3),
    ((3, 3), 2, 1024, (), (), 0.5),
    ((30, 40), 20, 600, (300,), (), 0.5),        # dense, 1 200 entries: the rows' change lists exceed the list capacity (row groups), 2 feature passes
    ((50, 60), 10, 128, (64,), (), 0.4),         # more entries per row than the default capacity
])
def test_fused_policy_matches_torch(obs_shape, A, emb, common, pol_layers, density):
    from qiskit_gym_b200.policy import FusedPolicy, pack_obs_bits
    from qiskit_gym_b200.search import BasicPolicy
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    pol = BasicPolicy(list(obs_shape), A, embedding_size=emb, common_layers=common, policy_layers=pol_layers).to(dev).eval()
    fused = FusedPolicy(pol, device=dev)
    rng = np.random.Generator(np.random.PCG64(9))
    for B in (1, 7, 8, 9, 1000):
        dense = torch.from_numpy((rng.random((B,) + tuple(obs_shape)) < density).astype(np.float32))
        if B > 2:
            dense[0] = 0.0                                          # an empty observation and a full one
            dense[1] = 1.0
        bits = pack_obs_bits(dense).to(dev)
        probs = torch.empty((B, A), dtype=torch.float32, device=dev)
        logits = torch.empty((B, A), dtype=torch.float32, device=dev)
        fused.forward_bits(bits, probs=probs, logits=logits)
        with torch.no_grad():
            # reference in float64 so the comparison is not limited by cuBLAS' own f32 rounding / TF32 settings
            ref_logits, _ = pol.double()(dense.to(dev).double())
            pol.float()
        ref_probs = torch.softmax(ref_logits, dim=-1)
        # tolerance: f32 accumulation over <= 1024-term sums
        assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=2e-5), (B, (logits.double() - ref_logits).abs().max())
        assert torch.allclose(probs.double(), ref_probs, rtol=1e-4, atol=1e-6)
        assert torch.allclose(probs.sum(dim=1), torch.ones(B, device=dev), atol=1e-5)
    if not pol_layers:
        # the value head rides along as one more output of the last layer (qg_policy_create_value): same logits, values of the module
        fv = FusedPolicy(pol, device=dev, with_value=True)
        values = torch.empty(B, dtype=torch.float32, device=dev)
        probs2 = torch.empty((B, A), dtype=torch.float32, device=dev)
        fv.forward_bits(bits, probs=probs2, values=values)
        with torch.no_grad():
            _, ref_v = pol.double()(dense.to(dev).double())
            pol.float()
        assert torch.allclose(values.double(), ref_v.reshape(-1), rtol=1e-4, atol=2e-5), (values.double() - ref_v.reshape(-1)).abs().max()
        assert torch.allclose(probs2.double(), ref_probs, rtol=1e-4, atol=1e-6)
    else:
        with pytest.raises(NotImplementedError):
            FusedPolicy(pol, device=dev, with_value=True)


@pytest.mark.gpu
def test_rollout_search_fused_backend_solves_shallow_targets():
    from qiskit_gym_b200.search import BasicPolicy, RolloutSearch
    kind, n, gs, kw = H.config_table()["C1_perm_grid3"]
    torch.manual_seed(0)
    pol = BasicPolicy([n, n], len(gs), embedding_size=64, common_layers=(32,))
    results = {}
    for backend in ("torch", "fused"):
        rs = RolloutSearch(kind, n, gs, pol, 2048, max_depth=6, policy_backend=backend, add_inverts=False)
        tgt = orc.OracleEnv(kind, n, gs, difficulty=2, add_inverts=False, add_perms=False)
        solved = 0
        for trial in range(4):
            tgt.reset(seed=trial, env_id=0)
            state = tgt.raw_state().astype(np.int64).tolist()
            res = rs.solve(state, deterministic=False, seed=trial)
            if res.actions is not None:
                solved += 1
                chk = orc.OracleEnv(kind, n, gs, add_inverts=False, add_perms=False)
                chk.set_state(state)
                for a in res.actions:
                    chk.step(a)
                assert chk.success()
        results[backend] = solved
        assert solved >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("name,deterministic,inverts", [("C1_perm_grid3", False, False), ("clifford3_allgates", False, False), ("lf5_line_swap", False, True),
                                                        ("pauli3_line", False, False), ("C5_perm27_heavyhex", True, False), ("C1_perm_grid3", False, True)])
def test_persistent_search_equals_two_kernel_search(name, deterministic, inverts):
    """qg_search_run (the whole search in one launch) takes exactly the decisions of the per-decision launches of
    qg_policy_forward_bits + qg_search_step_bits: same returns (f32 bits), same final states, same best rollout, same solution."""
    from qiskit_gym_b200.search import BasicPolicy, RolloutSearch
    kind, n, gs, kw = H.config_table()[name]
    kw = dict(kw)
    if kind != H.PAULI:
        kw["add_inverts"] = inverts
    R = 203
    torch.manual_seed(2)
    probe = orc.OracleEnv(kind, n, gs, add_perms=False, **kw)
    pol = BasicPolicy(probe.obs_shape(), len(gs), embedding_size=64, common_layers=(32,))
    tgt = orc.OracleEnv(kind, n, gs, difficulty=4, add_perms=False, **kw)
    tgt.reset(seed=9, env_id=0)
    if kind == H.PAULI:
        t = H.random_targets(kind, n, gs, 1, 3, scramble=3, num_rotations=2)
        state = t[0, : H.payload_lengths(kind, n, t)[0]].tolist()
    else:
        state = tgt.raw_state().astype(np.int64).tolist()
    out = {}
    for backend in ("fused", "persistent"):
        rs = RolloutSearch(kind, n, gs, pol, R, max_depth=10, policy_backend=backend, use_cuda_graph=False, **kw)
        res = rs.solve(state, deterministic=deterministic, seed=5, first_rollout_id=1000)
        out[backend] = (res.key, res.actions, rs.env.returns().cpu().numpy().view(np.uint32).copy(), [rs.env.get_state(b) for b in (0, 7, 8, 100, R - 1)],
                        rs.env.status()[3].cpu().numpy().copy())
    a, b = out["fused"], out["persistent"]
    assert a[0] == b[0] and a[1] == b[1]
    assert np.array_equal(a[2], b[2])
    assert all(np.array_equal(x, y) for x, y in zip(a[3], b[3]))
    assert np.array_equal(a[4], b[4])


def test_fixed_point_first_layer_is_at_least_as_accurate_as_f32():
    """The first layer's arithmetic (qg_policy.cu: weights -> int32 with the largest |weight| just below 2^30, int64 sums, one rounding to
    f32, f32 bias add), restated with numpy: against the exact (float64) sum its error is below that of a sequential float32 sum of the
    same terms, for trained-looking and for badly scaled weights."""
    rng = np.random.Generator(np.random.PCG64(5))
    for scale, obs, width, density in ((0.05, 729, 512, 0.04), (1.0, 256, 512, 0.5), (30.0, 500, 300, 0.3), (1e-3, 64, 64, 0.5)):
        w = (rng.standard_normal((obs, width)) * scale).astype(np.float32)
        w[0, 0] = np.float32(8.0 * scale)                                   # an outlier sets the fixed-point scale
        bias = (rng.standard_normal(width) * scale).astype(np.float32)
        mx = float(np.abs(w).max())
        shift = min(30 - int(np.frexp(mx)[1]), 120)
        q = np.rint(np.ldexp(w.astype(np.float64), shift)).astype(np.int64)
        assert np.abs(q).max() < 2**30
        err_fixed, err_f32 = [], []
        for _ in range(8):
            on = rng.random(obs) < density
            exact = w[on].astype(np.float64).sum(axis=0) + bias.astype(np.float64)
            acc = q[on].sum(axis=0)                                          # exact integers, any order
            fixed = (np.float32(acc.astype(np.float64)) * np.float32(np.ldexp(1.0, -shift))).astype(np.float32) + bias
            seq = np.zeros(width, np.float32)
            for row in w[on]:
                seq = (seq + row).astype(np.float32)
            seq = (seq + bias).astype(np.float32)
            err_fixed.append(np.abs(fixed.astype(np.float64) - exact).max())
            err_f32.append(np.abs(seq.astype(np.float64) - exact).max())
        assert max(err_fixed) <= max(err_f32) * 1.01 + 1e-12, (scale, max(err_fixed), max(err_f32))
        assert max(err_fixed) <= 2e-5 * max(1.0, mx * 8)
