"""Device tree search (csrc/qg_mcts.cu, mcts.MCTSSearch) against a CPU restatement of the same protocol built on the oracle's
envs: same descents, same created nodes, same visit counts (f32 bit for bit), same decisions.  The policy outputs the trees
consume are recorded on the GPU and replayed on the CPU, so the comparison is about the search and the env, not cuBLAS."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H
from tests.test_search import STREAM_SAMPLE, cpu_pick

f32 = np.float32


class CpuTree:
    """The protocol stated at the top of csrc/qg_mcts.cu, one tree."""

    def __init__(self, root_env, prior, A, c_puct):
        self.A, self.c = A, f32(c_puct)
        self.envs = [root_env]
        self.P = [prior.astype(f32)]
        self.N = [np.zeros(A, np.int64)]
        self.W = [np.zeros(A, f32)]
        self.child = [np.full(A, -1, np.int64)]
        self.reward = [f32(0)]
        self.final = [root_env.is_final()]

    def select(self):
        """-> (path, parent, action) with action = -1 when nothing is expanded"""
        path, node = [], 0
        if self.final[0]:
            return path, 0, -1
        while True:
            n = self.N[node]
            sq = np.sqrt(f32(n.sum() + 1)).astype(f32)
            q = np.where(n > 0, self.W[node] / np.maximum(n, 1).astype(f32), f32(0)).astype(f32)
            u = (((self.c * self.P[node]).astype(f32) * sq).astype(f32) / (n + 1).astype(f32)).astype(f32)
            a = int(np.argmax((q + u).astype(f32)))
            path.append((node, a))
            c = int(self.child[node][a])
            if c < 0:
                return path, node, a
            node = c
            if self.final[node]:
                return path, node, -1

    def expand(self, parent, a):
        env = self.envs[parent].clone()
        env.step(a)
        self.child[parent][a] = len(self.envs)
        self.envs.append(env)
        self.N.append(np.zeros(self.A, np.int64)); self.W.append(np.zeros(self.A, f32)); self.child.append(np.full(self.A, -1, np.int64))
        self.P.append(None)
        self.reward.append(f32(env.reward())); self.final.append(env.is_final())
        return len(self.envs) - 1

    def backup(self, path, leaf_value):
        G = f32(leaf_value)
        for node, a in reversed(path):
            G = f32(self.reward[int(self.child[node][a])] + G)
            self.N[node][a] += 1
            self.W[node][a] = f32(self.W[node][a] + G)

    def root_weights(self):
        n = self.N[0]
        tot = n.sum()
        return (n.astype(f32) / f32(tot)).astype(f32) if tot > 0 else np.zeros(self.A, f32)


@pytest.mark.gpu
@pytest.mark.parametrize("name,deterministic", [("C1_perm_grid3", False), ("clifford3_allgates", True), ("lf5_line_swap", False), ("pauli3_line", False)])
def test_tree_search_matches_cpu_restatement(name, deterministic):
    from qiskit_gym_b200.mcts import MCTSSearch
    from qiskit_gym_b200.search import BasicPolicy

    kind, n, gs, kw = H.config_table()[name]
    kw = dict(kw, add_perms=False)
    if kind != H.PAULI:
        kw["add_inverts"] = False
    R, S, decisions, seed, first, cp = 37, 12, 5, 21, 300, 1.5
    A = len(gs)
    tgt_env = orc.OracleEnv(kind, n, gs, difficulty=3, **kw)
    tgt_env.reset(seed=4, env_id=0)
    if kind == H.PAULI:
        t = H.random_targets(kind, n, gs, 1, 3, scramble=3, num_rotations=2)
        target = t[0, : H.payload_lengths(kind, n, t)[0]].tolist()
    else:
        target = tgt_env.raw_state().astype(np.int64).tolist()
    torch.manual_seed(1)
    ekw = dict(max_depth=6, **kw)
    probe = orc.OracleEnv(kind, n, gs, **ekw)
    pol = BasicPolicy(probe.obs_shape(), A, embedding_size=32, common_layers=(16,))
    ms = MCTSSearch(kind, n, gs, pol, R, S, C=cp, **ekw)
    rec = {}
    ms.hook = lambda kind_, d, s, p, v: rec.__setitem__((kind_, d, s), (p.cpu().numpy().copy(), None if v is None else v.cpu().numpy().copy()))
    ms.env.set_state(target)
    ms.env.search_begin(seed, first)
    refs = []
    for b in range(R):
        r = orc.OracleEnv(kind, n, gs, **ekw)
        r.set_state(target)
        refs.append(r)
    ticks = np.zeros(R, np.int64)
    chosen = torch.zeros(R, dtype=torch.int32, device=ms.env.device)
    expansions = 0
    for d in range(decisions):
        w_gpu = ms.decide(d).cpu().numpy()
        ms.env.search_step(ms.weights, deterministic=deterministic, obs=False, chosen=chosen)
        got = chosen.cpu().numpy()
        root_prior = rec[("root", d, -1)][0]
        for b, r in enumerate(refs):
            tree = CpuTree(r.clone(), root_prior[b], A, cp)
            for s in range(S):
                p, v = rec[("leaf", d, s)]
                path, parent, a = tree.select()
                if a >= 0:
                    new = tree.expand(parent, a)
                    tree.P[new] = p[b].astype(f32)
                    expansions += 1
                    tree.backup(path, f32(0) if tree.final[new] else v[b])
                else:
                    tree.backup(path, f32(0))
            w = tree.root_weights()
            assert np.array_equal(w.view(np.uint32), w_gpu[b].view(np.uint32)), (d, b, w, w_gpu[b])
            if r.is_final():
                assert got[b] == -1
                continue
            raw = orc.philox_draw(seed, first + b, int(ticks[b]), STREAM_SAMPLE)
            act = cpu_pick(w, raw, deterministic)
            assert got[b] == act, (d, b)
            r.step(act)
            ticks[b] += 1
    assert expansions > 2 * R            # trees really grew (not only root-final shortcuts)
    for b in (0, 5, R - 1):
        assert np.array_equal(ms.env.get_state(b), refs[b].raw_state())


@pytest.mark.gpu
def test_mcts_solve_finds_shallow_targets():
    from qiskit_gym_b200.mcts import MCTSSearch
    from qiskit_gym_b200.search import BasicPolicy
    kind, n, gs, kw = H.config_table()["C1_perm_grid3"]
    torch.manual_seed(0)
    pol = BasicPolicy([n, n], len(gs), embedding_size=64, common_layers=(32,))
    ms = MCTSSearch(kind, n, gs, pol, 64, 16, max_depth=6, add_inverts=False)
    tgt = orc.OracleEnv(kind, n, gs, difficulty=2, add_inverts=False, add_perms=False)
    solved = 0
    for trial in range(3):
        tgt.reset(seed=trial, env_id=0)
        state = tgt.raw_state().astype(np.int64).tolist()
        res = ms.solve(state, deterministic=False, seed=trial)
        if res.actions is not None:
            solved += 1
            chk = orc.OracleEnv(kind, n, gs, add_inverts=False, add_perms=False)
            chk.set_state(state)
            for a in res.actions:
                chk.step(a)
            assert chk.success()
    assert solved >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("sims", [12, 130])
def test_graph_replay_equals_eager_tree_search(sims):
    """The captured decision (one graph up to 128 simulations, root + per-simulation graphs beyond) leaves the same visit weights, bit
    for bit, as the launch-by-launch search; prints the time of both."""
    import time
    from qiskit_gym_b200.mcts import MCTSSearch
    from qiskit_gym_b200.search import BasicPolicy
    kind, n, gs, kw = H.config_table()["C2_lf8_line"]
    torch.manual_seed(1)
    pol = BasicPolicy([n, n], len(gs), embedding_size=64, common_layers=(32,))
    R = 96
    tarr = H.random_targets(kind, n, gs, R, 3, scramble=6)
    out = {}
    for graph in (False, True):
        ms = MCTSSearch(kind, n, gs, pol, R, sims, max_depth=16, add_inverts=False, use_cuda_graph=graph)
        ms.env.set_state(tarr)
        ms.env.search_begin(5, 0)
        w = []
        for d in range(3):
            w.append(ms.decide(d).clone())
            ms.env.search_step(ms.weights, deterministic=True, obs=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for d in range(3):
            ms.decide(d)
        torch.cuda.synchronize()
        out[graph] = (w, (time.perf_counter() - t0) / 3)
    for a, b in zip(out[False][0], out[True][0]):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    print(f"tree search decision, {R} rollouts x {sims} simulations: eager {out[False][1] * 1e3:.2f} ms, graph {out[True][1] * 1e3:.2f} ms")
