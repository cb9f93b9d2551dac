"""configs.py against the reference's own config classes: tests/golden/config_kats.json holds `to_json()` outputs produced by
importing /root/reference/src/qiskit_gym/rl/configs.py (tests/golden/make_golden.py::make_config_kats)."""
import json
import os

import pytest

from qiskit_gym_b200 import configs as C

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config_kats.json")))


def test_defaults_match_the_reference():
    assert C.PPOConfig().to_json() == KATS["ppo_default"]
    assert C.AlphaZeroConfig().to_json() == KATS["az_default"]
    assert C.BasicPolicyConfig().to_json() == KATS["basic_default"]
    assert C.Conv1dPolicyConfig().to_json() == KATS["conv_default"]
    assert C.PPOConfig().algorithm_cls == "twisterl.rl.PPO" and C.AlphaZeroConfig().algorithm_cls == "twisterl.rl.AZ"
    assert C.BasicPolicyConfig().policy_cls == "twisterl.nn.BasicPolicy" and C.Conv1dPolicyConfig().policy_cls == "twisterl.nn.Conv1dPolicy"
    assert set(C.ALGORITHMS) == {"PPO", "AZ"} and set(C.POLICIES) == {"BasicPolicy", "Conv1dPolicy"}


def test_custom_values_and_partial_from_json():
    ppo = C.PPOConfig(num_episodes=64, gae_lambda=0.9, lr=1e-3, diff_max=12, evals={"quick": C.EvalConfig(num_episodes=8)}, diff_metric="quick")
    assert ppo.to_json() == KATS["ppo_custom"]
    az = C.AlphaZeroConfig(num_mcts_searches=32, C=2.0, evals={"m": C.EvalConfig(num_mcts_searches=8)}, diff_metric="m")
    assert az.to_json() == KATS["az_custom"]
    assert C.PPOConfig.from_json({"collecting": {"num_episodes": 7}, "evals": {"x": {"num_searches": 3}}}).to_json() == KATS["ppo_from_partial"]
    assert C.BasicPolicyConfig(embedding_size=64, common_layers=[32, 16], value_layers=[8]).to_json() == KATS["basic_custom"]
    # round trips
    # (from_json seeds the evals with the defaults, like the reference's: identity only when the defaults are present)
    assert C.PPOConfig.from_json(KATS["ppo_default"]).to_json() == KATS["ppo_default"]
    assert C.AlphaZeroConfig.from_json(KATS["az_default"]).to_json() == KATS["az_default"]
    rt = C.PPOConfig.from_json(KATS["ppo_custom"]).to_json()
    assert {k: v for k, v in rt.items() if k != "evals"} == {k: v for k, v in KATS["ppo_custom"].items() if k != "evals"}
    assert rt["evals"]["quick"] == KATS["ppo_custom"]["evals"]["quick"] and "ppo_10" in rt["evals"]
    assert C.BasicPolicyConfig.from_json(KATS["basic_custom"]).to_json() == KATS["basic_custom"]
    assert C.PPOConfig().with_updates(lr=1e-2).lr == 1e-2


def test_validation_errors():
    for bad in (dict(num_episodes=0), dict(gae_lambda=1.5), dict(clip_ratio=0.0), dict(diff_metric="missing"), dict(diff_threshold=2.0)):
        with pytest.raises(ValueError):
            C.PPOConfig(**bad).validate()
    for bad in (dict(num_mcts_searches=0), dict(C=0.0), dict(max_expand_depth=0)):
        with pytest.raises(ValueError):
            C.AlphaZeroConfig(**bad).validate()
    with pytest.raises(ValueError):
        C.PPOConfig(evals={"ppo_deterministic": C.EvalConfig(num_searches=0)}).validate()
    with pytest.raises(ValueError):
        C.BasicPolicyConfig(common_layers=[0]).validate()
    with pytest.raises(ValueError):
        C.BasicPolicyConfig(embedding_size=0).validate()


def test_trainer_reads_config_objects():
    from qiskit_gym_b200 import ppo
    assert ppo.merged_config(C.PPOConfig(num_episodes=32).to_json())["collecting"]["num_episodes"] == 32
    assert ppo.merged_config(C.AlphaZeroConfig(num_mcts_searches=8).to_json(), "AZ")["collecting"]["num_mcts_searches"] == 8
