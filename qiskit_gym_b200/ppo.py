"""PPO over the on-device collector — what `RLSynthesis.learn()` runs (reference src/qiskit_gym/rl/synthesis.py:128-139
-> twisterl `PPO.learn`; hyper-parameters and their JSON schema from src/qiskit_gym/rl/configs.py:72-240).

The reference hands the env to twisterl, whose Rust collectors step `num_episodes` cloned envs on a rayon pool and whose
Python side runs the clipped-surrogate update.  Here the collection is `collector.RolloutCollector` (every env of one
`BatchedEnv` plays episodes back to back on the GPU, GAE on the device) and the update is plain PyTorch on the tensors
the collector left on the device; nothing crosses the PCIe bus during training except the logged scalars.

Config: the nested dict `PPOConfig.to_json()` writes (configs.py:205-240) —
    collecting {num_cores (ignored: the batch is the parallelism), num_episodes, lambda, gamma}
    training   {num_epochs, vf_coef, ent_coef, clip_ratio, normalize_advantage}
    learning   {diff_threshold, diff_max, diff_metric}      optimizer {lr}
    evals      {name: {num_episodes, deterministic, num_searches, num_mcts_searches, num_cores, C}}
    logging    {log_freq, checkpoint_freq}
Missing keys take the reference's defaults.  twisterl itself is not in the reference tree, so the loop below follows its
documented semantics (configs.py docstrings), not its code: one iteration = collect >= num_episodes episodes at the current
difficulty, `num_epochs` full-batch updates, evaluate, and raise the difficulty by one when `evals[diff_metric]` reaches
`diff_threshold` (up to `diff_max`).
"""
from __future__ import annotations

import json
import os
import time

import torch

from .collector import RolloutCollector, decision_seed
from .engine import BatchedEnv

_DEFAULTS = {
    "collecting": {"num_cores": 32, "num_episodes": 1024, "lambda": 0.995, "gamma": 0.995},
    "training": {"num_epochs": 10, "vf_coef": 0.8, "ent_coef": 0.01, "clip_ratio": 0.1, "normalize_advantage": False},
    "learning": {"diff_threshold": 0.85, "diff_max": 256, "diff_metric": "ppo_deterministic"},
    "optimizer": {"lr": 3e-4},
    "evals": {"ppo_deterministic": {"num_episodes": 100, "deterministic": True, "num_searches": 1, "num_mcts_searches": 0, "num_cores": 32, "C": 1.41},
              "ppo_10": {"num_episodes": 100, "deterministic": False, "num_searches": 10, "num_mcts_searches": 0, "num_cores": 32, "C": 1.41}},
    "logging": {"log_freq": 1, "checkpoint_freq": 10},
}


def merged_config(cfg: dict | None) -> dict:
    """The reference's defaults (configs.py:133-166) overlaid with `cfg`; validates like `PPOConfig.validate` (168-196)."""
    out = {k: dict(v) for k, v in _DEFAULTS.items()}
    out["evals"] = {k: dict(v) for k, v in _DEFAULTS["evals"].items()}
    for sec, val in (cfg or {}).items():
        if sec == "evals":
            out["evals"] = {name: {**_DEFAULTS["evals"]["ppo_deterministic"], **dict(ev)} for name, ev in dict(val).items()}
        elif sec in out:
            out[sec].update(dict(val))
    c, t, l = out["collecting"], out["training"], out["learning"]
    if c["num_episodes"] <= 0:
        raise ValueError("num_episodes must be > 0")
    if not 0.0 <= c["lambda"] <= 1.0:
        raise ValueError("gae_lambda must be in [0, 1]")
    if not 0.0 <= c["gamma"] <= 1.0:
        raise ValueError("gamma must be in [0, 1]")
    if t["num_epochs"] <= 0:
        raise ValueError("num_epochs must be > 0")
    if t["clip_ratio"] <= 0:
        raise ValueError("clip_ratio must be > 0")
    if not 0.0 <= l["diff_threshold"] <= 1.0:
        raise ValueError("diff_threshold must be in [0, 1]")
    if l["diff_max"] < 1:
        raise ValueError("diff_max must be >= 1")
    if l["diff_metric"] not in out["evals"]:
        raise ValueError(f"diff_metric '{l['diff_metric']}' not found in evals: {list(out['evals'])}")
    for name, ev in out["evals"].items():
        if ev["num_episodes"] <= 0 or ev["num_searches"] <= 0 or ev["num_mcts_searches"] < 0:
            raise ValueError(f"Invalid eval '{name}'")
    return out


class PPO:
    """`PPO(env_spec, policy, config).learn(...)`; env_spec = (env kind, num_qubits, gateset, BatchedEnv keyword arguments)."""

    def __init__(self, env_kind: int, num_qubits: int, gateset, policy: torch.nn.Module, config: dict | None = None, device=None,
                 seed: int = 0, use_twists: bool = True, minibatch_size: int | None = None, **env_kwargs):
        self.cfg = merged_config(config)
        self.spec = (env_kind, num_qubits, list(gateset), dict(env_kwargs))
        B = int(self.cfg["collecting"]["num_episodes"])
        self.env = BatchedEnv(env_kind, num_qubits, gateset, B, device=device, **env_kwargs)
        self.device = self.env.device
        self.policy = policy.to(self.device)
        self.collector = RolloutCollector(self.env, self.policy, gamma=self.cfg["collecting"]["gamma"], lam=self.cfg["collecting"]["lambda"],
                                          use_twists=use_twists, seed=seed)
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=float(self.cfg["optimizer"]["lr"]))
        self.seed = int(seed)
        self.minibatch_size = minibatch_size
        self._eval_envs = {}
        self.iteration = 0
        self.history = []

    # ---- difficulty (Env::set_difficulty on every env the trainer owns) -------------------------------------------
    @property
    def difficulty(self) -> int:
        return int(self.env.difficulty)

    @difficulty.setter
    def difficulty(self, d: int):
        self.env.difficulty = int(d)
        for e in self._eval_envs.values():
            e.difficulty = int(d)

    def _episode_steps(self) -> int:
        """An episode lasts at most depth = min(depth_slope * difficulty, max_depth) decisions (e.g. clifford.rs:306-319)."""
        c = self.env.cfg
        return int(max(1, min(int(c.depth_slope) * self.difficulty, int(c.max_depth))))

    # ---- one PPO iteration ---------------------------------------------------------------------------------------------
    def update(self, ro) -> dict:
        """`num_epochs` passes of the clipped-surrogate update over the valid decisions of a Rollout."""
        t = self.cfg["training"]
        valid = ro.valid.reshape(-1)
        idx = valid.nonzero(as_tuple=False).squeeze(1)
        obs = ro.obs.reshape((-1,) + tuple(ro.obs.shape[2:]))[idx]
        act = ro.policy_actions.reshape(-1)[idx]
        logp_old = ro.logp.reshape(-1)[idx]
        adv = ro.advantages.reshape(-1)[idx]
        ret = ro.returns.reshape(-1)[idx]
        if t["normalize_advantage"] and adv.numel() > 1:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        n = int(idx.numel())
        mb = n if not self.minibatch_size else min(int(self.minibatch_size), n)
        clip = float(t["clip_ratio"])
        self.policy.train()
        stats = {}
        for _ in range(int(t["num_epochs"])):
            perm = torch.randperm(n, device=self.device) if mb < n else None
            for s in range(0, n, mb):
                sel = slice(s, s + mb) if perm is None else perm[s:s + mb]
                logits, value = self.policy(obs[sel])
                logp_all = torch.log_softmax(logits.float(), dim=-1)
                logp = logp_all.gather(1, act[sel][:, None]).squeeze(1)
                ratio = torch.exp(logp - logp_old[sel])
                a = adv[sel]
                pi_loss = -torch.min(ratio * a, torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * a).mean()
                v_loss = torch.nn.functional.mse_loss(value.float().reshape(-1), ret[sel])
                entropy = -(logp_all.exp() * logp_all).sum(-1).mean()
                loss = pi_loss + float(t["vf_coef"]) * v_loss - float(t["ent_coef"]) * entropy
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                self.opt.step()
                stats = {"loss": loss.detach(), "pi_loss": pi_loss.detach(), "v_loss": v_loss.detach(), "entropy": entropy.detach(),
                         "clip_frac": ((ratio - 1.0).abs() > clip).float().mean().detach()}
        self.policy.eval()
        return {k: float(v.item()) for k, v in stats.items()} | {"samples": n}

    # ---- evaluation ----------------------------------------------------------------------------------------------------
    def evaluate(self, name: str) -> float:
        """Success rate of `evals[name]`: num_episodes fresh targets at the current difficulty, each tried with `num_searches`
        rollouts (greedy or sampled); an episode counts when any of its rollouts ends in success (configs.py:25-35)."""
        ev = self.cfg["evals"][name]
        if int(ev.get("num_mcts_searches", 0)):
            raise NotImplementedError("evals with num_mcts_searches > 0: use RLSynthesis.solve(num_mcts_searches=...)")
        E = int(ev["num_episodes"])
        env = self._eval_envs.get(E)
        if env is None:
            kind, n, gs, kw = self.spec
            env = BatchedEnv(kind, n, gs, E, device=self.device.index, **kw)
            self._eval_envs[E] = env
        env.difficulty = self.difficulty
        base = decision_seed(self.seed ^ 0x5EED5EED, 1_000_003 * (self.iteration + 1))
        env.reset(seed=base)
        env.snapshot()
        solved = torch.zeros(E, dtype=torch.bool, device=self.device)
        steps = self._episode_steps()
        for s in range(int(ev["num_searches"])):
            if s:
                env.restore()
            for k in range(steps):
                with torch.no_grad():
                    logits, _ = self.policy(env.observe())
                    probs = torch.softmax(logits.float(), dim=-1).contiguous()
                env.collect_step(probs, decision_seed(base, (s + 1) * 4099 + k), deterministic=bool(ev["deterministic"]), obs=False)
            solved |= env.status()[2]
        return float(solved.float().mean().item())

    # ---- the loop ------------------------------------------------------------------------------------------------------
    def learn(self, initial_difficulty: int = 1, num_iterations: int = int(1e10), tb_path: str | None = None, log=None, stop_at_max: bool = False):
        """rl/synthesis.py:128-139.  Returns the history (one dict per iteration).  `tb_path`: directory for `metrics.jsonl` and
        checkpoints (`checkpoint_<iteration>.pt`, a plain state_dict like the reference's examples/models)."""
        L, lg = self.cfg["learning"], self.cfg["logging"]
        self.difficulty = int(initial_difficulty)
        fh = None
        if tb_path:
            os.makedirs(tb_path, exist_ok=True)
            fh = open(os.path.join(tb_path, "metrics.jsonl"), "a")
        try:
            for _ in range(int(num_iterations)):
                t0 = time.time()
                ro = self.collector.collect(self._episode_steps())
                episodes, success = ro.episode_stats()
                rec = {"iteration": self.iteration, "difficulty": self.difficulty, "episodes": episodes, "collect_success": success,
                       "mean_reward": float(ro.rewards[ro.valid].mean().item()) if bool(ro.valid.any()) else 0.0}
                rec.update(self.update(ro))
                if self.iteration % max(1, int(lg["log_freq"])) == 0:
                    for name in self.cfg["evals"]:
                        rec[f"eval/{name}"] = self.evaluate(name)
                    metric = rec[f"eval/{L['diff_metric']}"]
                    if metric >= float(L["diff_threshold"]) and self.difficulty < int(L["diff_max"]):
                        self.difficulty = self.difficulty + 1
                rec["seconds"] = time.time() - t0
                self.history.append(rec)
                if fh:
                    fh.write(json.dumps(rec) + "\n")
                    fh.flush()
                    if (self.iteration + 1) % max(1, int(lg["checkpoint_freq"])) == 0:
                        torch.save(self.policy.state_dict(), os.path.join(tb_path, f"checkpoint_{self.iteration + 1}.pt"))
                if log:
                    log(rec)
                self.iteration += 1
                if stop_at_max and self.difficulty >= int(L["diff_max"]) and rec.get(f"eval/{L['diff_metric']}", 0.0) >= float(L["diff_threshold"]):
                    break
        finally:
            if fh:
                fh.close()
        return self.history
