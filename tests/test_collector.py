"""Rollout-collector pieces (qg_reset_select, qg_collect_step, qg_gae, qg_twist_gather, collector.RolloutCollector):
the env side against a CPU re-enactment with the oracle (same Philox streams, same f32 sums); the twist convention
against the symmetry it is meant to express (CPU only)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H
from tests.test_search import STREAM_SAMPLE, cpu_pick


def cpu_gae(reward, value, done, valid, gamma, lam):
    """qg_gae restated with numpy float32 scalars (every operation rounded on its own, same order)."""
    T, B = reward.shape
    f = np.float32
    adv = np.zeros((T, B), f)
    ret = np.zeros((T, B), f)
    gl = f(f(gamma) * f(lam))
    for b in range(B):
        nv, na = f(value[T, b]), f(0)
        for t in range(T - 1, -1, -1):
            v = f(value[t, b])
            if valid is not None and not valid[t, b]:
                na, nv = f(0), v
                continue
            nd = f(0) if done[t, b] else f(1)
            delta = f(f(f(reward[t, b]) + f(f(f(gamma) * nv) * nd)) - v)
            a = f(delta + f(f(gl * nd) * na))
            adv[t, b], ret[t, b] = a, f(a + v)
            na, nv = a, v
    return adv, ret


def dense(obs_idx, size):
    o = np.zeros(size, np.float32)
    o[obs_idx] = 1.0
    return o


# ---------------------------------------------------------------------------------------------------- CPU: convention
@pytest.mark.parametrize("name", ["lf5_line_swap", "clifford3_allgates", "C1_perm_grid3"])
def test_twist_convention_commutes_with_dynamics(name):
    """Entry i of the observation moves to obs_perms[k][i]; env action g is action act_perms[k][g] of the twisted frame
    (symmetry.rs:297-361).  With that reading, twisting then stepping equals stepping then twisting — the property twists
    exist for.  (The opposite reading fails this test for non-involutive symmetries.)"""
    kind, n, gs, kw = H.config_table()[name]
    kw = dict(kw, add_inverts=False, add_perms=True)
    env = orc.OracleEnv(kind, n, gs, difficulty=12, **kw)
    obs_perms, act_perms = env.twists()
    assert len(obs_perms) >= 2
    size = env.obs_shape()[0] * env.obs_shape()[1]
    rng = np.random.Generator(np.random.PCG64(3))
    for trial in range(6):
        env.reset(seed=trial, env_id=7)
        base = env.raw_state().astype(np.int64)
        for k in range(len(obs_perms)):
            op, ap = np.asarray(obs_perms[k]), np.asarray(act_perms[k])
            a = orc.OracleEnv(kind, n, gs, **kw)
            a.set_state(base.tolist())
            tw = np.zeros(size, np.float32)
            tw[op] = dense(a.observe(), size)                      # entry i -> obs_perms[k][i]
            # the state whose observation is the twisted observation
            b = orc.OracleEnv(kind, n, gs, **kw)
            if kind == H.PERM:
                b.set_state(np.argmax(tw.reshape(n, n), axis=1).tolist())
            else:
                b.set_state(tw.astype(np.int64).tolist())
            assert np.array_equal(dense(b.observe(), size), tw)
            for _ in range(5):
                g = int(rng.integers(len(gs)))
                a.step(g)
                b.step(int(ap[g]))
                tw2 = np.zeros(size, np.float32)
                tw2[op] = dense(a.observe(), size)
                assert np.array_equal(dense(b.observe(), size), tw2), (name, k, g)
                assert a.reward() == b.reward() and a.is_final() == b.is_final()


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gae_matches_cpu():
    from qiskit_gym_b200.collector import gae
    rng = np.random.Generator(np.random.PCG64(11))
    for T, B in ((1, 1), (7, 33), (40, 300)):
        reward = rng.normal(size=(T, B)).astype(np.float32)
        value = rng.normal(size=(T + 1, B)).astype(np.float32)
        done = rng.random((T, B)) < 0.2
        valid = rng.random((T, B)) < 0.9
        dev = torch.device("cuda", 0)
        for vd in (None, valid):
            adv, ret = gae(torch.from_numpy(reward).to(dev), torch.from_numpy(value).to(dev), torch.from_numpy(done).to(dev), 0.995, 0.95,
                           None if vd is None else torch.from_numpy(vd).to(dev))
            ra, rr = cpu_gae(reward, value, done, vd, 0.995, 0.95)
            assert np.array_equal(adv.cpu().numpy().view(np.uint32), ra.view(np.uint32))
            assert np.array_equal(ret.cpu().numpy().view(np.uint32), rr.view(np.uint32))


@pytest.mark.gpu
def test_twist_gather_matches_numpy():
    from qiskit_gym_b200.collector import twist_gather
    rng = np.random.Generator(np.random.PCG64(12))
    dev = torch.device("cuda", 0)
    for B, L, K in ((1, 1, 1), (37, 81, 8), (1000, 256, 5), (513, 12, 3)):
        table = np.stack([rng.permutation(L) for _ in range(K)]).astype(np.int32)
        idx = rng.integers(K, size=B).astype(np.int32)
        src = rng.random((B, L)).astype(np.float32)
        out = twist_gather(torch.from_numpy(src).to(dev), torch.from_numpy(table).to(dev), torch.from_numpy(idx).to(dev))
        assert np.array_equal(out.cpu().numpy(), np.take_along_axis(src, table[idx].astype(np.int64), axis=1))
        out0 = twist_gather(torch.from_numpy(src).to(dev), torch.from_numpy(table).to(dev), None)
        assert np.array_equal(out0.cpu().numpy(), src[:, table[0]])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C1_perm_grid3", "clifford3_allgates", "lf5_line_swap", "pauli3_line"])
def test_reset_select_only_touches_final_envs(name):
    from qiskit_gym_b200 import BatchedEnv
    kind, n, gs, kw = H.config_table()[name]
    kw = dict(kw, add_perms=False)
    if kind != H.PAULI:
        kw["add_inverts"] = False
    B = 77
    env = BatchedEnv(kind, n, gs, B, difficulty=2, max_depth=4, depth_slope=1, **kw)
    env.reset(seed=5, first_env_id=100)
    rng = np.random.Generator(np.random.PCG64(2))
    refs = []
    for b in range(B):
        r = orc.OracleEnv(kind, n, gs, difficulty=2, max_depth=4, depth_slope=1, **kw)
        r.reset(seed=5, env_id=100 + b)
        refs.append(r)
    for t in range(6):
        acts = rng.integers(len(gs), size=B).astype(np.int32)
        env.step(torch.from_numpy(acts).to(env.device))
        for b, r in enumerate(refs):
            r.step(int(acts[b]))
        if t % 2 == 1:
            env.reset_select(seed=1000 + t, first_env_id=100)
            n_final = 0
            for b, r in enumerate(refs):
                if r.is_final():
                    n_final += 1
                    r.reset(seed=1000 + t, env_id=100 + b)
            assert n_final > 0
        for b in (0, 1, 5, 33, 76):
            assert np.array_equal(env.get_state(b), refs[b].raw_state()), (t, b)
        _, done, succ, depth = env.status()
        assert done.cpu().numpy().tolist() == [r.is_final() for r in refs]
        assert depth.cpu().numpy().tolist() == [r.depth() for r in refs]
    # explicit selection
    sel = np.zeros(B, np.uint8)
    sel[[3, 40]] = 1
    before = [env.get_state(b) for b in (2, 3, 40)]
    env.reset_select(seed=77, first_env_id=100, select=torch.from_numpy(sel).to(env.device))
    for b in (3, 40):
        refs[b].reset(seed=77, env_id=100 + b)
        assert np.array_equal(env.get_state(b), refs[b].raw_state())
    assert np.array_equal(env.get_state(2), before[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name,twists", [("C1_perm_grid3", True), ("clifford3_allgates", True), ("lf5_line_swap", False), ("pauli3_line", False)])
def test_collector_matches_cpu_reenactment(name, twists):
    from qiskit_gym_b200 import BatchedEnv
    from qiskit_gym_b200.collector import RolloutCollector, decision_seed
    from qiskit_gym_b200.search import BasicPolicy

    kind, n, gs, kw = H.config_table()[name]
    kw = dict(kw, add_perms=twists)              # Perm / LF / Clifford: add_perms only switches twists() on
    if kind != H.PAULI:
        kw["add_inverts"] = False
    B, T, seed, first = 70, 24, 1234, 500
    ekw = dict(difficulty=2, max_depth=5, depth_slope=2, **kw)
    env = BatchedEnv(kind, n, gs, B, **ekw)
    torch.manual_seed(0)
    pol = BasicPolicy(env.obs_shape(), len(gs), embedding_size=32, common_layers=(16,))
    col = RolloutCollector(env, pol, gamma=0.99, lam=0.9, use_twists=twists, seed=seed, first_env_id=first)
    assert (col.num_twists > 1) == twists
    weights = []
    col.hook = lambda t, w: weights.append(w.detach().cpu().numpy().copy())
    ro = col.collect(T)
    refs = [orc.OracleEnv(kind, n, gs, **ekw) for _ in range(B)]
    ticks = np.zeros(B, np.int64)
    acts = ro.actions.cpu().numpy()
    rew = ro.rewards.cpu().numpy()
    dones = ro.dones.cpu().numpy()
    succ = ro.successes.cpu().numpy()
    obs = ro.obs.cpu().numpy().reshape(T, B, -1)
    size = obs.shape[-1]
    tw = None if ro.twist is None else ro.twist.cpu().numpy()
    obs_perms = env.twists()[0] if twists else None
    n_resets = n_invalid = 0
    for t in range(T):
        s = decision_seed(seed, t)
        for b, r in enumerate(refs):
            if r.is_final():                       # constructor state is final too
                r.reset(seed=s, env_id=first + b)
                ticks[b] = 0
                n_resets += 1
            if kind != H.PAULI:                    # (PauliNetwork observations need its internal perm pick; covered by the parity tests)
                want = dense(r.observe(), size)
                if twists:
                    tmp = np.zeros(size, np.float32)
                    tmp[np.asarray(obs_perms[tw[t, b]])] = want
                    want = tmp
                assert np.array_equal(obs[t, b], want), (t, b)
            if r.is_final():                       # reset produced a solved state: not stepped
                assert acts[t, b] == -1
                n_invalid += 1
                continue
            raw = orc.philox_draw(s, first + b, int(ticks[b]), STREAM_SAMPLE)
            a = cpu_pick(weights[t][b], raw, False)
            assert acts[t, b] == a, (t, b)
            r.step(a)
            ticks[b] += 1
            assert np.float32(r.reward()).view(np.uint32) == rew[t, b].view(np.uint32)
            assert bool(dones[t, b]) == r.is_final() and bool(succ[t, b]) == r.success()
    assert n_resets > B                            # episodes really ended and restarted inside the rollout
    valid = acts >= 0
    assert np.array_equal(ro.valid.cpu().numpy(), valid)
    ra, rr = cpu_gae(rew, ro.values.cpu().numpy(), dones, valid, 0.99, 0.9)
    assert np.array_equal(ro.advantages.cpu().numpy().view(np.uint32), ra.view(np.uint32))
    assert np.array_equal(ro.returns.cpu().numpy().view(np.uint32), rr.view(np.uint32))
    # log-probabilities are those of the sampled actions under the recorded weights
    w = np.stack(weights)
    lp = np.log(np.take_along_axis(w, np.maximum(acts, 0)[..., None].astype(np.int64), axis=2)[..., 0])
    assert np.allclose(ro.logp.cpu().numpy()[valid], lp[valid], rtol=1e-5, atol=1e-6)
    # a second collect continues the same envs and the same seed sequence
    ro2 = col.collect(3)
    assert ro2.actions.shape == (3, B) and col.counter == T + 3
