"""`RLSynthesis` over the device-resident search (reference src/qiskit_gym/rl/synthesis.py:32-147).

The reference builds a twisterl algorithm object around the raw env and calls `algorithm.solve(state,
deterministic, num_searches, num_mcts_searches, C, max_expand_depth)` (112-126).  Here `solve` is
`search.RolloutSearch`: `num_searches` policy-guided rollouts stepped together on the GPU, the best one reduced on
the device (and across ranks).  Config files and checkpoints are the reference's own formats: the `.json` written by
`RLSynthesis.save` (79-93) and the policy `state_dict` `.pt` (examples/models/*), so models trained with the
reference load unchanged.

`num_mcts_searches > 0` runs the device tree search of mcts.py.  `learn()` runs PPO over collector.RolloutCollector or AlphaZero
self-play over the device tree search (ppo.py), chosen by `algorithm_cls` like the reference does.
"""
from __future__ import annotations

import json

import torch

from .search import BasicPolicy, RolloutSearch
from .specs import SynthSpec

_ENV_KINDS = {"PermutationEnv": 0, "LinearFunctionEnv": 1, "CliffordEnv": 2, "PauliNetworkEnv": 3}


def gate_list_to_circuit(gate_list, num_qubits=None):
    """rl/synthesis.py:141-147 (needs Qiskit)."""
    if num_qubits is None:
        num_qubits = max(max(gate_args) for _, gate_args in gate_list) + 1
    from .wire import gates_to_circuit
    return gates_to_circuit(gate_list, num_qubits)


class RLSynthesis:
    def __init__(self, env, rl_config: dict | None, model_config: dict | None, model_path: str | None = None, device=None,
                 policy_cls: str = "twisterl.nn.BasicPolicy", algorithm_cls: str = "twisterl.rl.PPO"):
        self.env = env
        self.env_config = env.to_json()
        # config objects (configs.PPOConfig / AlphaZeroConfig / BasicPolicyConfig, as in the reference) or dicts in their JSON schema
        if hasattr(rl_config, "to_json"):
            algorithm_cls, rl_config = rl_config.algorithm_cls, rl_config.to_json()
        if hasattr(model_config, "to_json"):
            policy_cls, model_config = model_config.policy_cls, model_config.to_json()
        self.rl_config = dict(rl_config or {})
        self.model_config = dict(model_config or {})
        self.policy_cls = policy_cls
        self.algorithm_cls = algorithm_cls
        if policy_cls.split(".")[-1] != "BasicPolicy":
            raise NotImplementedError(f"policy class {policy_cls} is not supported (BasicPolicy only)")
        self.device = device
        self.policy = self._init_policy(model_path)
        self._searches = {}

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_config_json(cls, config_path, model_path=None, device=None):
        """rl/synthesis.py:54-77: `{env_cls, env, policy_cls, policy, algorithm_cls, algorithm}`."""
        full = json.load(open(config_path))
        env = SynthSpec.from_json(full["env_cls"], full["env"])      # Qiskit-free problem spec (specs.py); a reference *Gym object works too
        return cls(env, full.get("algorithm"), full.get("policy"), model_path, device=device,
                   policy_cls=full.get("policy_cls", "twisterl.nn.BasicPolicy"), algorithm_cls=full.get("algorithm_cls", "twisterl.rl.PPO"))

    def _init_policy(self, model_path):
        """rl/synthesis.py:95-110: policy(obs_shape, num_actions, **model_config); checkpoint = plain state_dict."""
        mc = self.model_config
        pol = BasicPolicy(self.env.obs_shape(), self.env.num_actions(), embedding_size=mc.get("embedding_size", 512),
                          common_layers=tuple(mc.get("common_layers", (256,))), policy_layers=tuple(mc.get("policy_layers", ())),
                          value_layers=tuple(mc.get("value_layers", ())))
        if model_path is not None:
            pol.load_state_dict(torch.load(model_path, map_location="cpu", weights_only=True))
        return pol.eval()

    def to_json(self):
        return {"env_cls": f"qiskit_gym.envs.synthesis.{self.env.cls_name}", "env": self.env_config, "policy_cls": self.policy_cls,
                "policy": self.model_config, "algorithm_cls": self.algorithm_cls, "algorithm": self.rl_config}

    def save(self, config_path, model_path=None):
        with open(config_path, "w") as f:
            json.dump(self.to_json(), f, indent=2)
        if model_path is not None:
            with open(model_path, "wb") as f:
                torch.save(self.policy.state_dict(), f)

    # ---- synthesis ------------------------------------------------------------------------------------
    def _search(self, num_searches: int) -> RolloutSearch:
        from .policy import weights_version
        rs = self._searches.get(num_searches)
        if rs is not None and getattr(rs, "fused", None) is not None and rs.fused.version != weights_version(self.policy):
            rs = None                                    # the cached search captured older weights (load_state_dict after the first synth)
        if rs is None:
            cfg = dict(self.env_config)
            kind = _ENV_KINDS[self.env.cls_name]
            kw = {k: v for k, v in cfg.items() if k not in ("num_qubits", "gateset", "max_depth", "add_perms")}

            def make(backend):
                return RolloutSearch(kind, cfg["num_qubits"], cfg["gateset"], self.policy, num_searches, device=self.device,
                                     max_depth=cfg.get("max_depth", 128), add_perms=False, policy_backend=backend, **kw)
            try:
                rs = make("persistent")
                rs.check_supported()
            except NotImplementedError:
                # outside the one-launch search's limits (Permutation wider than 64 qubits, a layer wider than 1024, policy + env larger
                # than one SM's shared memory): the PyTorch policy on dense observations has none of them
                rs = make("torch")
            self._searches[num_searches] = rs
        return rs

    def _mcts(self, num_searches: int, num_mcts_searches: int, C: float):
        from .mcts import MCTSSearch
        key = ("mcts", num_searches, num_mcts_searches, C)
        ms = self._searches.get(key)
        if ms is None:
            cfg = dict(self.env_config)
            kind = _ENV_KINDS[self.env.cls_name]
            kw = {k: v for k, v in cfg.items() if k not in ("num_qubits", "gateset", "max_depth", "add_perms")}
            ms = MCTSSearch(kind, cfg["num_qubits"], cfg["gateset"], self.policy, num_searches, num_mcts_searches, C=C, device=self.device,
                            max_depth=cfg.get("max_depth", 128), add_perms=False, **kw)
            self._searches[key] = ms
        return ms

    def solve(self, state, deterministic: bool = False, num_searches: int = 100, num_mcts_searches: int = 0, C: float = 2 ** 0.5,
              max_expand_depth: int = 1, seed: int = 0):
        """twisterl `Algorithm.solve`: the action list of the best successful rollout, or None."""
        if num_mcts_searches:
            if max_expand_depth != 1:
                raise NotImplementedError("max_expand_depth != 1 is not supported by the device tree search (mcts.py)")
            return self._mcts(int(num_searches), int(num_mcts_searches), float(C)).solve(state, deterministic=deterministic, seed=seed).actions
        return self._search(int(num_searches)).solve(state, deterministic=deterministic, seed=seed).actions

    def synth(self, input, deterministic: bool = False, num_searches: int = 100, num_mcts_searches: int = 0, C: float = 2 ** 0.5,
              max_expand_depth: int = 1, seed: int = 0):
        """rl/synthesis.py:112-126.  Returns the synthesised circuit (QuantumCircuit with Qiskit, else a gate list) or None."""
        state = self.env.get_state(input)
        actions = self.solve(state, deterministic, num_searches, num_mcts_searches, C, max_expand_depth, seed=seed)
        if actions is not None:
            return self.env.build_circuit_from_solution(actions, input)

    def learn(self, initial_difficulty=1, num_iterations=int(1e10), tb_path=None, log=None, seed: int = 0):
        """rl/synthesis.py:128-139: PPO (`twisterl.rl.PPO`) or AlphaZero (`twisterl.rl.AZ`) on the device (ppo.py).  The trained
        policy is `self.policy` (searches built before the call are dropped so that `synth` uses the new weights)."""
        from . import ppo
        algo = self.algorithm_cls.split(".")[-1]
        if algo not in ("PPO", "AZ"):
            raise NotImplementedError(f"algorithm class {self.algorithm_cls} is not supported (twisterl.rl.PPO, twisterl.rl.AZ)")
        cfg = dict(self.env_config)
        kind = _ENV_KINDS[self.env.cls_name]
        kw = {k: v for k, v in cfg.items() if k not in ("num_qubits", "gateset")}
        trainer = (ppo.PPO if algo == "PPO" else ppo.AlphaZero)(kind, cfg["num_qubits"], cfg["gateset"], self.policy, self.rl_config, device=self.device, seed=seed, **kw)
        if hasattr(self.env, "difficulty"):
            self.env.difficulty = initial_difficulty                # rl/synthesis.py:129 (a reference *Gym object; a SynthSpec owns no env)
        try:
            return trainer.learn(initial_difficulty=initial_difficulty, num_iterations=num_iterations, tb_path=tb_path, log=log)
        except KeyboardInterrupt:
            # rl/synthesis.py:137-139: stopping a run by hand keeps the policy trained so far (metrics.jsonl under tb_path holds the history)
            return getattr(trainer, "history", None)
        finally:
            self.policy = trainer.policy.eval()
            self._searches = {}
