"""Synth search pieces: GPU sampling/stepping/return accumulation and best-rollout reduction against a CPU
re-enactment with the oracle (same Philox draws, same f32 sums); cross-rank reduction logic under gloo."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H

STREAM_SAMPLE = 4


def cpu_pick(weights, raw, deterministic):
    """The selection rule of qg_search_step, restated: first arg-max, or inverse CDF over a sequential f32 cumsum."""
    w = weights.astype(np.float32)
    if deterministic:
        return int(np.argmax(w))
    cum = np.cumsum(w, dtype=np.float32)
    total = cum[-1]
    if not total > 0:
        return int((int(raw) * len(w)) >> 32)
    target = np.float32(np.float32(raw >> 8) * np.float32(1.0 / 16777216.0)) * total
    hit = np.nonzero(cum > target)[0]
    return int(hit[0]) if hit.size else len(w) - 1


def order_key(success, ret, gid):
    u = int(np.float32(ret).view(np.uint32))
    u = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
    return (int(success) << 62) | (u << 30) | ((0x3FFFFFFF - gid) & 0x3FFFFFFF)


@pytest.mark.gpu
@pytest.mark.parametrize("name,deterministic", [("C1_perm_grid3", False), ("C1_perm_grid3", True), ("clifford3_allgates", False), ("pauli3_line", False), ("lf5_line_swap", False)])
def test_search_step_matches_cpu_reenactment(name, deterministic):
    from qiskit_gym_b200 import BatchedEnv

    kind, n, gs, kw = H.config_table()[name]
    B, T, seed, first = 150, 12, 99, 4000
    rng = np.random.Generator(np.random.PCG64(5))
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = False
    # a target a few gates away from identity so that some rollouts succeed early and stop
    tgt_env = orc.OracleEnv(kind, n, gs, difficulty=2, add_perms=False, **pk)
    tgt_env.reset(seed=1, env_id=0)
    if kind == H.PAULI:
        target = H.random_targets(kind, n, gs, 1, 3, scramble=2, num_rotations=2)
        target = target[0, : H.payload_lengths(kind, n, target)[0]].tolist()
    else:
        target = tgt_env.raw_state().astype(np.int64).tolist()
    A = len(gs)
    env = BatchedEnv(kind, n, gs, B, max_depth=8, add_perms=False, **pk)
    env.set_state(target)
    env.search_begin(seed, first)
    refs = []
    for b in range(B):
        r = orc.OracleEnv(kind, n, gs, max_depth=8, add_perms=False, **pk)
        r.set_state(target)
        refs.append(r)
    rets = np.zeros(B, dtype=np.float32)
    ticks = np.zeros(B, dtype=np.int64)
    chosen = torch.zeros(B, dtype=torch.int32, device=env.device)
    nact = torch.zeros(1, dtype=torch.int32, device=env.device)
    for t in range(T):
        w = rng.random((B, A)).astype(np.float32)
        w[rng.random((B, A)) < 0.3] = 0.0
        env.search_step(torch.from_numpy(w).to(env.device), deterministic=deterministic, chosen=chosen, num_active=nact)
        got = chosen.cpu().numpy()
        active = 0
        for b, r in enumerate(refs):
            if r.is_final():
                assert got[b] == -1
                continue
            active += 1
            raw = orc.philox_draw(seed, first + b, int(ticks[b]), STREAM_SAMPLE)
            a = cpu_pick(w[b], raw, deterministic)
            assert got[b] == a, (t, b)
            r.step(a)
            rets[b] = np.float32(rets[b] + np.float32(r.reward()))
            ticks[b] += 1
        assert int(nact.item()) == active
    assert np.array_equal(env.returns().cpu().numpy().view(np.uint32), rets.view(np.uint32))
    keys = [order_key(r.success(), rets[b], first + b) for b, r in enumerate(refs)]
    key, idx = env.search_best()
    assert key == max(keys) and idx == int(np.argmax(keys))
    assert env.solution(idx) == refs[idx].solution()


@pytest.mark.gpu
def test_rollout_search_solves_shallow_targets():
    """End-to-end RolloutSearch (policy MLP + CUDA graph): a uniform policy finds depth-2 permutation targets and the
    returned action list, applied to the target, gives the identity."""
    from qiskit_gym_b200.search import BasicPolicy, RolloutSearch

    kind, n, gs, kw = H.config_table()["C1_perm_grid3"]
    torch.manual_seed(0)
    pol = BasicPolicy([n, n], len(gs), embedding_size=64, common_layers=(32,))
    for use_graph in (False, True):
        rs = RolloutSearch(kind, n, gs, pol, 2048, max_depth=6, use_cuda_graph=use_graph, add_inverts=False)
        tgt = orc.OracleEnv(kind, n, gs, difficulty=2, add_inverts=False, add_perms=False)
        solved = 0
        for trial in range(4):
            tgt.reset(seed=trial, env_id=0)
            state = tgt.raw_state().astype(np.int64).tolist()
            res = rs.solve(state, deterministic=False, seed=trial)
            assert res.rollouts == 2048
            if res.actions is not None:
                solved += 1
                chk = orc.OracleEnv(kind, n, gs, add_inverts=False, add_perms=False)
                chk.set_state(state)
                for a in res.actions:
                    chk.step(a)
                assert chk.success() and res.success
        assert solved >= 3
    # an unreachable budget returns None (rl/synthesis.py:125)
    rs = RolloutSearch(kind, n, gs, pol, 64, max_depth=1, add_inverts=False)
    far = [8, 7, 6, 5, 4, 3, 2, 1, 0]
    assert rs.solve(far, seed=1).actions is None


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from qiskit_gym_b200.search import decode_key, reduce_best
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    # case 1: rank 1 owns the best (successful) rollout
    keys = [order_key(False, -0.3, 5), order_key(True, 0.7, 1000 + 17)]
    sols = [[1, 2, 3], [9, 8, 7, 6]]
    out.append(reduce_best(keys[rank], sols[rank]))
    # case 2: equal success and return -> lowest rollout id wins (rank 0)
    keys = [order_key(True, 0.5, 3), order_key(True, 0.5, 1003)]
    out.append(reduce_best(keys[rank], sols[rank]))
    # case 3: nobody ran any rollout
    out.append(reduce_best(0, None))
    # case 4: higher return beats lower id
    keys = [order_key(True, 0.25, 0), order_key(True, 0.5, 1500)]
    out.append(reduce_best(keys[rank], sols[rank]))
    q.put((rank, out, [decode_key(k) for k, _ in out]))
    dist.destroy_process_group()


def test_cross_rank_best_rollout_reduction_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, dec in res:
        assert out[0][1] == [9, 8, 7, 6] and dec[0] == (True, 1017)
        assert out[1][1] == [1, 2, 3] and dec[1] == (True, 3)
        assert out[2] == (0, None)
        assert out[3][1] == [9, 8, 7, 6] and dec[3] == (True, 1500)
    assert res[0][1] == res[1][1]          # every rank ends with the same winner
