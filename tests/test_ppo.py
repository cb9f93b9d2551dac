"""`RLSynthesis.learn()` / ppo.PPO (reference rl/synthesis.py:128-139, rl/configs.py:72-240): the config schema on the CPU, and on
the GPU a short training run — a policy trained from scratch on the engine must climb the difficulty curriculum and then
synthesise targets it could not solve before training."""
import json

import numpy as np
import pytest
import torch

from qiskit_gym_b200 import ppo


def test_config_defaults_are_the_reference_defaults():
    c = ppo.merged_config(None)
    # rl/configs.py:133-166
    assert c["collecting"] == {"num_cores": 32, "num_episodes": 1024, "lambda": 0.995, "gamma": 0.995}
    assert c["training"] == {"num_epochs": 10, "vf_coef": 0.8, "ent_coef": 0.01, "clip_ratio": 0.1, "normalize_advantage": False}
    assert c["learning"] == {"diff_threshold": 0.85, "diff_max": 256, "diff_metric": "ppo_deterministic"}
    assert c["optimizer"] == {"lr": 3e-4}
    assert set(c["evals"]) == {"ppo_deterministic", "ppo_10"}
    assert c["evals"]["ppo_10"]["num_searches"] == 10 and c["evals"]["ppo_10"]["deterministic"] is False
    assert c["logging"] == {"log_freq": 1, "checkpoint_freq": 10}


def test_config_overlay_and_validation():
    c = ppo.merged_config({"collecting": {"num_episodes": 64}, "evals": {"quick": {"num_episodes": 8}}, "learning": {"diff_metric": "quick"}})
    assert c["collecting"]["num_episodes"] == 64 and c["collecting"]["gamma"] == 0.995
    assert c["evals"] == {"quick": {"num_episodes": 8, "deterministic": True, "num_searches": 1, "num_mcts_searches": 0, "num_cores": 32, "C": 1.41}}
    for bad in ({"collecting": {"num_episodes": 0}}, {"collecting": {"lambda": 1.5}}, {"training": {"clip_ratio": 0.0}},
                {"learning": {"diff_metric": "missing"}}, {"learning": {"diff_threshold": 2.0}}, {"evals": {"ppo_deterministic": {"num_searches": 0}}}):
        with pytest.raises(ValueError):
            ppo.merged_config(bad)


def test_alphazero_config_defaults_and_validation():
    c = ppo.merged_config(None, "AZ")
    # rl/configs.py:325-360
    assert c["collecting"] == {"num_cores": 32, "num_episodes": 128, "num_mcts_searches": 1000, "C": 1.41, "max_expand_depth": 1}
    assert c["training"] == {"num_epochs": 10} and c["learning"]["diff_metric"] == "mcts_100"
    assert c["evals"]["mcts_100"]["num_mcts_searches"] == 100 and c["evals"]["mcts_100"]["deterministic"] is True
    for bad in ({"collecting": {"num_mcts_searches": 0}}, {"collecting": {"C": 0.0}}, {"collecting": {"max_expand_depth": 0}},
                {"learning": {"diff_metric": "nope"}}):
        with pytest.raises(ValueError):
            ppo.merged_config(bad, "AZ")


def test_reference_config_files_parse():
    """The `algorithm` section of the reference's own example configs (tests/golden/models/*.json) goes through unchanged."""
    import glob
    import os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "models", "*.json")))
    assert files
    for f in files:
        full = json.load(open(f))
        c = ppo.merged_config(full.get("algorithm"))
        assert c["learning"]["diff_metric"] in c["evals"]


@pytest.mark.gpu
def test_ppo_learns_permutation_line4(tmp_path):
    from qiskit_gym_b200 import gyms
    from qiskit_gym_b200.rl import RLSynthesis

    torch.manual_seed(0)
    env = gyms.PermutationGym.from_coupling_map([(0, 1), (1, 2), (2, 3)], difficulty=1, depth_slope=2, max_depth=32)
    cfg = {"collecting": {"num_episodes": 512}, "training": {"num_epochs": 4, "ent_coef": 0.01}, "optimizer": {"lr": 2e-3},
           "learning": {"diff_threshold": 0.85, "diff_max": 8, "diff_metric": "ppo_deterministic"},
           "evals": {"ppo_deterministic": {"num_episodes": 128}}, "logging": {"log_freq": 1, "checkpoint_freq": 20}}
    rls = RLSynthesis(env, cfg, {"embedding_size": 64, "common_layers": [64]}, device=0)
    targets = [list(np.random.Generator(np.random.PCG64(s)).permutation(4)) for s in range(24)]
    targets = [t for t in targets if t != [0, 1, 2, 3]]
    before = sum(rls.synth(t, deterministic=True, num_searches=1) is not None for t in targets)
    hist = rls.learn(initial_difficulty=1, num_iterations=40, tb_path=str(tmp_path))
    assert len(hist) == 40 and hist[-1]["difficulty"] >= 4, [(h["difficulty"], round(h["eval/ppo_deterministic"], 2)) for h in hist]
    assert (tmp_path / "metrics.jsonl").exists() and (tmp_path / "checkpoint_20.pt").exists()
    after = 0
    for t in targets:
        circ = rls.synth(t, deterministic=True, num_searches=1)
        if circ is not None:
            after += 1
            gates = circ if isinstance(circ, list) else [(i.operation.name.upper(), [circ.find_bit(q).index for q in i.qubits]) for i in circ.data]
            assert all(g[0] == "SWAP" for g in gates)
    assert after >= max(before + 5, int(0.8 * len(targets))), (before, after, len(targets))
    # the saved checkpoint is a plain state_dict that a fresh RLSynthesis loads (rl/synthesis.py:95-110)
    rls.save(str(tmp_path / "cfg.json"), str(tmp_path / "model.pt"))
    again = RLSynthesis.from_config_json(str(tmp_path / "cfg.json"), str(tmp_path / "model.pt"), device=0)
    assert (again.synth(targets[0], deterministic=True, num_searches=1) is None) == (rls.synth(targets[0], deterministic=True, num_searches=1) is None)


@pytest.mark.gpu
def test_alphazero_learns_permutation_line4():
    """twisterl.rl.AZ: self-play with the device tree search; the curriculum must move and the tree-search eval must beat chance."""
    from qiskit_gym_b200 import gyms
    from qiskit_gym_b200.rl import RLSynthesis

    torch.manual_seed(0)
    env = gyms.PermutationGym.from_coupling_map([(0, 1), (1, 2), (2, 3)], difficulty=1, depth_slope=2, max_depth=32)
    cfg = {"collecting": {"num_episodes": 256, "num_mcts_searches": 16, "C": 1.41}, "training": {"num_epochs": 4}, "optimizer": {"lr": 2e-3},
           "learning": {"diff_threshold": 0.85, "diff_max": 6, "diff_metric": "mcts_16"},
           "evals": {"ppo_deterministic": {"num_episodes": 64}, "mcts_16": {"num_episodes": 64, "num_mcts_searches": 16}}}
    rls = RLSynthesis(env, cfg, {"embedding_size": 64, "common_layers": [64]}, device=0, algorithm_cls="twisterl.rl.AZ")
    hist = rls.learn(initial_difficulty=1, num_iterations=25)
    trace = [(h["difficulty"], round(h["eval/mcts_16"], 2), round(h["eval/ppo_deterministic"], 2)) for h in hist]
    assert hist[-1]["difficulty"] >= 3, trace
    assert max(h["eval/ppo_deterministic"] for h in hist[-5:]) >= 0.5, trace
    assert all(np.isfinite(h["loss"]) for h in hist)
