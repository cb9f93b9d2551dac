"""Synth-time rollout search on the device (the `num_searches` policy-guided rollouts of
`RLSynthesis.synth`, reference src/qiskit_gym/rl/synthesis.py:112-126 -> twisterl `Algorithm.solve`).

Per decision step, for all rollouts at once:   observation (device tensor, written by the previous fused step)
-> PyTorch policy MLP -> softmax -> `qg_search_step` (masked arg-max / Philox sampling + fused env step +
return accumulation + action log).  One iteration is captured in a CUDA graph and replayed; the best rollout is
reduced on the GPU (`qg_search_best`) and, across ranks, with one int64 MAX all-reduce plus a broadcast of the
winning action list.

twisterl is not in the reference tree, so its exact rollout protocol is *unpinned* (SURVEY.md §8c): what is
fixed here is  (a) final rollouts stop stepping, (b) best = (success, sum of rewards, lowest rollout id),
(c) `None` when no rollout succeeds (rl/synthesis.py:125).
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import torch

from .engine import BatchedEnv

KEY_SUCCESS_BIT = 62
KEY_ID_MASK = 0x3FFFFFFF


class BasicPolicy(torch.nn.Module):
    """MLP with the parameter names/shapes of twisterl's `BasicPolicy` as saved in the reference's
    examples/models/*.pt: embeddings.{weight[E,obs],bias}, common.{i}.{weight,bias}, action.0.*, value.0.*.
    (Activation placement is recalled, not verifiable here: ReLU after the embedding and after each common layer.)"""

    def __init__(self, obs_shape, num_actions, embedding_size=512, common_layers=(256,), policy_layers=(), value_layers=()):
        super().__init__()
        obs = 1
        for d in obs_shape:
            obs *= int(d)
        self.embeddings = torch.nn.Linear(obs, embedding_size)
        layers, last = [], embedding_size
        for h in common_layers:
            layers += [torch.nn.Linear(last, h), torch.nn.ReLU()]
            last = h
        self.common = torch.nn.Sequential(*layers)

        def head(hidden, out):
            ls, l = [], last
            for h in hidden:
                ls += [torch.nn.Linear(l, h), torch.nn.ReLU()]
                l = h
            ls.append(torch.nn.Linear(l, out))
            return torch.nn.Sequential(*ls)

        self.action = head(tuple(policy_layers), num_actions)
        self.value = head(tuple(value_layers), 1)

    def forward(self, obs):
        x = torch.relu(self.embeddings(obs.reshape(obs.shape[0], -1)))
        x = self.common(x)
        return self.action(x), self.value(x)


@dataclass
class SearchResult:
    actions: list | None          # Env::solution of the best rollout, or None if none succeeded
    key: int                      # packed key of the best rollout over all ranks
    rollout_id: int               # global id of the best rollout
    success: bool
    iterations: int               # decision steps executed
    seconds: float                # wall time of the whole search incl. the reduction
    rollouts: int                 # rollouts run over all ranks


def decode_key(key: int):
    return bool((key >> KEY_SUCCESS_BIT) & 1), KEY_ID_MASK - (key & KEY_ID_MASK)


def reduce_best(local_key: int, local_solution, group=None, device=None):
    """Cross-rank best-rollout reduction in two collectives: an all-gather of (packed key, solution length) — the owner of the winning
    rollout is the rank holding the largest key; keys are unique, they embed the global rollout id — then the owner broadcasts its action
    list.  Works with any torch.distributed backend (NCCL on GPUs, gloo in tests)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_key, local_solution
    dev = device if device is not None else torch.device("cpu")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sol = list(local_solution) if local_solution is not None else []
    mine = torch.tensor([local_key, len(sol)], dtype=torch.int64, device=dev)
    allk = torch.empty((world, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allk, mine, group=group) if dev.type == "cuda" else dist.all_gather(list(allk.unbind(0)), mine, group=group)
    table = allk.tolist()
    best = max(k for k, _ in table)
    if best == 0:
        return 0, None
    src = next(r for r, (k, _) in enumerate(table) if k == best)       # (the lowest rank on the impossible tie)
    n = int(table[src][1])
    buf = torch.zeros(max(n, 1), dtype=torch.int64, device=dev)
    if rank == src and n:
        buf[:n] = torch.tensor(sol, dtype=torch.int64)
    dist.broadcast(buf, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return best, [int(v) for v in buf[:n].tolist()]


class RolloutSearch:
    """Device-resident `solve`: owns a BatchedEnv of `num_rollouts` rollouts on one GPU."""

    def __init__(self, env_kind, num_qubits, gateset, policy: torch.nn.Module, num_rollouts: int, device=None,
                 max_depth: int = 128, use_cuda_graph: bool = True, policy_backend: str = "persistent", **env_kwargs):
        """policy_backend: "persistent" = the whole search in ONE kernel launch (qg_search_run: every CTA loops policy -> sample ->
        step for its 8 rollouts); "fused" = two launches per decision (policy.FusedPolicy + qg_search_step_bits) in a CUDA graph;
        "torch" = dense f32 observations + the PyTorch module (cuBLAS GEMMs, softmax).  "persistent" and "fused" take identical
        decisions (same kernels' device code, same bits)."""
        env_kwargs.setdefault("add_perms", False)
        self.env = BatchedEnv(env_kind, num_qubits, gateset, num_rollouts, device=device, max_depth=max_depth, **env_kwargs)
        self.policy = policy.to(self.env.device).eval()
        assert policy_backend in ("persistent", "fused", "torch")
        self.backend = policy_backend
        if policy_backend in ("fused", "persistent"):
            from .policy import FusedPolicy
            self.fused = FusedPolicy(self.policy, device=self.env.device)
            self.obs_bits = self.env.new_obs_bits()
        self.max_depth = max_depth
        self.B = num_rollouts
        dev = self.env.device
        self.probs = torch.zeros((self.B, self.env.num_actions()), dtype=torch.float32, device=dev)
        self.num_active = torch.zeros(1, dtype=torch.int32, device=dev)
        self.use_graph = use_cuda_graph
        self._graphs = {}
        self._stream = torch.cuda.Stream(device=dev)
        self._decisions = torch.zeros((self.B + 7) // 8, dtype=torch.int32, device=dev)

    def check_supported(self):
        """Raises NotImplementedError (QG_ERR_UNSUPPORTED) now rather than in the first solve() when the one-launch search cannot run
        this env / policy pair: a zero-decision qg_search_run goes through all of its checks without launching anything."""
        if self.backend == "persistent":
            self.env.search_run(self.fused, self.obs_bits, self.probs, 0, deterministic=True, decisions=self._decisions)

    def _iteration(self, deterministic):
        if self.backend == "fused":
            self.fused.forward_bits(self.obs_bits, probs=self.probs)
            self.env.search_step_bits(self.probs, self.obs_bits, deterministic=deterministic, num_active=self.num_active)
            return
        with torch.no_grad():
            logits, _ = self.policy(self.env.obs)
            torch.softmax(logits.float(), dim=-1, out=self.probs)
        self.env.search_step(self.probs, deterministic=deterministic, obs=True, num_active=self.num_active)

    def _observe(self):
        if self.backend in ("fused", "persistent"):
            self.env.observe_bits(self.obs_bits)
        else:
            self.env.observe()

    def _graph(self, deterministic):
        g = self._graphs.get(deterministic)
        if g is None:
            # warm-up outside capture (cuBLAS workspaces, lazy module init), on a scratch copy of the state
            self.env.snapshot()
            for _ in range(2):
                self._iteration(deterministic)
            self.env.restore()
            torch.cuda.current_stream(self.env.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._iteration(deterministic)
            self.env.restore()
            self._graphs[deterministic] = g
        return g

    def _local_best(self, comm):
        """(key, solution or None) of this rank's rollouts — or, with an ncclComm_t of the C ABI (engine.nccl_comm_create), of ALL ranks'
        rollouts: qg_search_finish does the cross-GPU reduction itself (one all-gather + on-GPU pick), no torch collective."""
        env = self.env
        if comm is not None:
            key, ok, rid, owner, sol = env.search_finish(comm)
            return key, sol, True
        key, idx = env.search_best()
        ok, rid = decode_key(key)
        return key, (env.solution(idx) if (ok and idx >= 0) else None), False

    def solve(self, state, deterministic: bool = False, seed: int = 0, first_rollout_id: int = 0, check_every: int = 8,
              group=None, comm=None) -> SearchResult:
        """group: torch.distributed group for the Python-side cross-rank reduction (reduce_best); comm: an ncclComm_t handle for the
        C-side one (qg_search_finish); with neither, and torch.distributed initialised, the default group is used."""
        env = self.env
        t0 = time.perf_counter()
        if self.backend == "persistent":
            with torch.cuda.stream(self._stream):
                env.set_state(state)
                env.search_begin(seed, first_rollout_id)
                self._observe()
                env.search_run(self.fused, self.obs_bits, self.probs, self.max_depth, deterministic=deterministic, decisions=self._decisions)
                key, sol, reduced = self._local_best(comm)
                its = int(self._decisions.max().item())
            return self._finish(key, sol, its, t0, group, reduced)
        with torch.cuda.stream(self._stream):
            env.set_state(state)                         # broadcast: every rollout starts from the target
            if self.use_graph:
                g = self._graph(deterministic)
                env.set_state(state)
            env.search_begin(seed, first_rollout_id)
            self._observe()
            its = 0
            while its < self.max_depth:
                if self.use_graph:
                    g.replay()
                else:
                    self._iteration(deterministic)
                its += 1
                if its % check_every == 0 and int(self.num_active.item()) == 0:
                    break
            key, sol, reduced = self._local_best(comm)
        return self._finish(key, sol, its, t0, group, reduced)

    def _finish(self, key, sol, its, t0, group, reduced=False):
        env = self.env
        ok, rid = decode_key(key)
        world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size(group)
                if not reduced:
                    key, sol = reduce_best(key, sol, group=group, device=env.device if dist.get_backend(group) == "nccl" else None)
                ok, rid = decode_key(key)
        except ImportError:
            pass
        return SearchResult(actions=sol if ok else None, key=key, rollout_id=rid, success=ok, iterations=its,
                            seconds=time.perf_counter() - t0, rollouts=self.B * world)
