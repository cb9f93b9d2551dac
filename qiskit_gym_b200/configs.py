"""Config objects with the reference's names, fields, defaults and JSON schema (src/qiskit_gym/rl/configs.py:20-705):
`EvalConfig`, `PPOConfig`, `AlphaZeroConfig`, `BasicPolicyConfig`, `Conv1dPolicyConfig`, and the `ALGORITHMS` / `POLICIES`
registries `RLSynthesis.from_config_json` looks classes up in.  `RLSynthesis(env, PPOConfig(...), BasicPolicyConfig(...))`
works like in the reference; plain dicts in the nested schema are accepted as well.

One declarative table per class (field -> JSON section / key) drives `to_json`, `from_json` and `with_updates`; validation is
the trainer's own (ppo.merged_config), so a config object and the equivalent dict are checked by the same code.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields, replace
from typing import Any, Dict, List, Mapping


@dataclass
class EvalConfig:
    """configs.py:20-69.  deterministic: greedy instead of sampled; num_searches: whole-episode rollouts, best kept;
    num_mcts_searches: tree-search simulations per decision; C: exploration constant."""
    num_episodes: int = 100
    deterministic: bool = True
    num_searches: int = 1
    num_mcts_searches: int = 0
    num_cores: int = 32
    C: float = 1.41

    def validate(self) -> None:
        for name, ok in (("num_episodes", self.num_episodes > 0), ("num_searches", self.num_searches > 0),
                         ("num_mcts_searches", self.num_mcts_searches >= 0), ("num_cores", self.num_cores > 0), ("C", self.C > 0)):
            if not ok:
                raise ValueError(f"EvalConfig.{name} out of range")

    @classmethod
    def from_partial(cls, data: Mapping[str, Any] | None) -> "EvalConfig":
        d = dict(data or {})
        base = cls()
        return cls(**{f.name: type(getattr(base, f.name))(d.get(f.name, getattr(base, f.name))) for f in fields(cls)})


def _default_evals(with_mcts: bool) -> Dict[str, EvalConfig]:
    ev = {"ppo_deterministic": EvalConfig(), "ppo_10": EvalConfig(deterministic=False, num_searches=10)}
    if with_mcts:
        ev["mcts_100"] = EvalConfig(deterministic=True, num_searches=1, num_mcts_searches=100)
    return ev


class _AlgorithmConfig:
    """to_json / from_json / validate / with_updates from the class's `_SCHEMA`: {section: {json key: field name}}."""
    _SCHEMA: Dict[str, Dict[str, str]] = {}
    _ALGO = "PPO"

    def to_json(self) -> dict:
        self.validate()
        out = {sec: {key: getattr(self, name) for key, name in keys.items()} for sec, keys in self._SCHEMA.items()}
        out["evals"] = {k: dict(vars(v)) for k, v in self.evals.items()}
        order = ("collecting", "training", "learning", "optimizer", "evals", "logging")
        return {k: out[k] for k in order}

    def validate(self) -> None:
        from .ppo import merged_config
        raw = {sec: {key: getattr(self, name) for key, name in keys.items()} for sec, keys in self._SCHEMA.items()}
        raw["evals"] = {k: dict(vars(v)) for k, v in self.evals.items()}
        for name, ev in self.evals.items():
            try:
                ev.validate()
            except Exception as e:
                raise ValueError(f"Invalid eval '{name}': {e}") from e
        merged_config(raw, self._ALGO)

    def with_updates(self, **kwargs):
        return replace(self, **kwargs)

    @classmethod
    def from_json(cls, data: Mapping[str, Any]):
        """Nested schema (`collecting / training / ...`); unknown keys are ignored, missing ones take the defaults; evals named
        in `data` are added to (or override) the default evals."""
        base = cls()
        kw = {}
        for sec, keys in cls._SCHEMA.items():
            for key, name in keys.items():
                kw[name] = dict(data.get(sec, {})).get(key, getattr(base, name))
        evals = dict(base.evals)
        for name, partial in dict(data.get("evals", {})).items():
            evals[name] = EvalConfig.from_partial(partial)
        obj = cls(**kw, evals=evals, algorithm_cls=data.get("algorithm_cls", base.algorithm_cls))
        obj.validate()
        return obj


@dataclass
class PPOConfig(_AlgorithmConfig):
    """configs.py:72-166."""
    num_cores: int = 32
    num_episodes: int = 1024
    gae_lambda: float = 0.995
    gamma: float = 0.995
    num_epochs: int = 10
    vf_coef: float = 0.8
    ent_coef: float = 0.01
    clip_ratio: float = 0.1
    normalize_advantage: bool = False
    lr: float = 3e-4
    diff_threshold: float = 0.85
    diff_max: int = 256
    diff_metric: str = "ppo_deterministic"
    evals: Dict[str, EvalConfig] = field(default_factory=lambda: _default_evals(False))
    log_freq: int = 1
    checkpoint_freq: int = 10
    algorithm_cls: str = "twisterl.rl.PPO"

    _ALGO = "PPO"
    _SCHEMA = {
        "collecting": {"num_cores": "num_cores", "num_episodes": "num_episodes", "lambda": "gae_lambda", "gamma": "gamma"},
        "training": {"num_epochs": "num_epochs", "vf_coef": "vf_coef", "ent_coef": "ent_coef", "clip_ratio": "clip_ratio",
                     "normalize_advantage": "normalize_advantage"},
        "learning": {"diff_threshold": "diff_threshold", "diff_max": "diff_max", "diff_metric": "diff_metric"},
        "optimizer": {"lr": "lr"},
        "logging": {"log_freq": "log_freq", "checkpoint_freq": "checkpoint_freq"},
    }


@dataclass
class AlphaZeroConfig(_AlgorithmConfig):
    """configs.py:295-360."""
    num_cores: int = 32
    num_episodes: int = 128
    num_mcts_searches: int = 1000
    C: float = 1.41
    max_expand_depth: int = 1
    num_epochs: int = 10
    lr: float = 3e-4
    diff_threshold: float = 0.85
    diff_max: int = 256
    diff_metric: str = "mcts_100"
    evals: Dict[str, EvalConfig] = field(default_factory=lambda: _default_evals(True))
    log_freq: int = 1
    checkpoint_freq: int = 10
    algorithm_cls: str = "twisterl.rl.AZ"

    _ALGO = "AZ"
    _SCHEMA = {
        "collecting": {"num_cores": "num_cores", "num_episodes": "num_episodes", "num_mcts_searches": "num_mcts_searches", "C": "C",
                       "max_expand_depth": "max_expand_depth"},
        "training": {"num_epochs": "num_epochs"},
        "learning": {"diff_threshold": "diff_threshold", "diff_max": "diff_max", "diff_metric": "diff_metric"},
        "optimizer": {"lr": "lr"},
        "logging": {"log_freq": "log_freq", "checkpoint_freq": "checkpoint_freq"},
    }


ALGORITHMS = {"PPO": PPOConfig, "AZ": AlphaZeroConfig}


def _check_layers(layers, name):
    if not isinstance(layers, list):
        raise ValueError(f"{name} must be a list of ints (got {type(layers).__name__}).")
    if any((not isinstance(x, int)) or x < 1 for x in layers):
        raise ValueError(f"Every entry in {name} must be an int >= 1 (got {layers}).")


class _PolicyConfig:
    def validate(self) -> None:
        if self.embedding_size < 1:
            raise ValueError("embedding_size must be >= 1.")
        for name in ("common_layers", "policy_layers", "value_layers"):
            _check_layers(getattr(self, name), name)

    def with_updates(self, **kwargs):
        return replace(self, **kwargs)

    def to_json(self) -> dict:
        self.validate()
        return {f.name: (list(getattr(self, f.name)) if isinstance(getattr(self, f.name), list) else getattr(self, f.name))
                for f in fields(self) if f.name != "policy_cls"}

    @classmethod
    def from_json(cls, data: Mapping[str, Any]):
        base = cls()
        kw = {}
        for f in fields(cls):
            v = data.get(f.name, getattr(base, f.name))
            kw[f.name] = list(v) if isinstance(getattr(base, f.name), list) else type(getattr(base, f.name))(v)
        obj = cls(**kw)
        obj.validate()
        return obj


@dataclass
class BasicPolicyConfig(_PolicyConfig):
    """configs.py:531-605: MLP torso + policy / value heads."""
    embedding_size: int = 512
    common_layers: List[int] = field(default_factory=lambda: [256])
    policy_layers: List[int] = field(default_factory=list)
    value_layers: List[int] = field(default_factory=list)
    policy_cls: str = "twisterl.nn.BasicPolicy"


@dataclass
class Conv1dPolicyConfig(_PolicyConfig):
    """configs.py:611-700.  The config object only: the Conv1dPolicy module itself lives in twisterl (not in the reference tree), and
    `RLSynthesis` here builds BasicPolicy networks only."""
    conv_dim: int = 1
    embedding_size: int = 1260
    common_layers: List[int] = field(default_factory=lambda: [256])
    policy_layers: List[int] = field(default_factory=list)
    value_layers: List[int] = field(default_factory=list)
    policy_cls: str = "twisterl.nn.Conv1dPolicy"


POLICIES = {"BasicPolicy": BasicPolicyConfig, "Conv1dPolicy": Conv1dPolicyConfig}
