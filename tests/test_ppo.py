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
    from qiskit_gym_b200.specs import SynthSpec
    from qiskit_gym_b200.rl import RLSynthesis

    torch.manual_seed(0)
    env = SynthSpec.from_coupling_map("PermutationEnv", [(0, 1), (1, 2), (2, 3)], difficulty=1, depth_slope=2, max_depth=32)
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
    from qiskit_gym_b200.specs import SynthSpec
    from qiskit_gym_b200.rl import RLSynthesis

    torch.manual_seed(0)
    env = SynthSpec.from_coupling_map("PermutationEnv", [(0, 1), (1, 2), (2, 3)], difficulty=1, depth_slope=2, max_depth=32)
    cfg = {"collecting": {"num_episodes": 256, "num_mcts_searches": 16, "C": 1.41}, "training": {"num_epochs": 4}, "optimizer": {"lr": 2e-3},
           "learning": {"diff_threshold": 0.85, "diff_max": 6, "diff_metric": "mcts_16"},
           "evals": {"ppo_deterministic": {"num_episodes": 64}, "mcts_16": {"num_episodes": 64, "num_mcts_searches": 16}}}
    rls = RLSynthesis(env, cfg, {"embedding_size": 64, "common_layers": [64]}, device=0, algorithm_cls="twisterl.rl.AZ")
    hist = rls.learn(initial_difficulty=1, num_iterations=25)
    trace = [(h["difficulty"], round(h["eval/mcts_16"], 2), round(h["eval/ppo_deterministic"], 2)) for h in hist]
    assert hist[-1]["difficulty"] >= 3, trace
    assert all(np.isfinite(h["loss"]) for h in hist)
    # the network itself (no tree search at eval time) must have learned from the self-play samples: this fails when the training
    # observations are not the ones the tree's root evaluated (a constant observation gives a state-independent policy)
    assert max(h["eval/ppo_deterministic"] for h in hist) >= 0.5, trace
    # the heads are being fitted: the value loss drops quickly; the policy's cross entropy against the (soft: 16 simulations over 3
    # actions) visit distributions creeps down from ln 3
    assert hist[-1]["v_loss"] < 0.6 * hist[0]["v_loss"], [round(h["v_loss"], 3) for h in hist]
    assert np.mean([h["pi_loss"] for h in hist[-5:]]) < hist[0]["pi_loss"], [round(h["pi_loss"], 3) for h in hist]


def _dp_worker(rank, world, port, q):
    """Two CPU replicas under gloo: equal start, different data, equal weights after synchronised steps."""
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                          # replicas start different ...
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ppo.broadcast_parameters(net)                          # ... and are made equal
    start = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(7 + rank)            # every rank sees its own batch
    x, y = torch.randn(8 + 4 * rank, 6, generator=g), torch.randn(8 + 4 * rank, 3, generator=g)
    steps = ppo.agree_min(3 + rank)                        # rank 0 would make 3 steps, rank 1 four: both make 3
    local_grads = []
    for _ in range(steps):
        opt.zero_grad()
        loss = ((net(x) - y) ** 2).mean()
        loss.backward()
        local_grads.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone())
        ppo.sync_gradients(list(net.parameters()))
        opt.step()
    end = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    synced = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    q.put((rank, start.tolist(), end.tolist(), steps, ppo.mean_over_ranks(float(rank + 1)), local_grads[-1].tolist(), synced.tolist()))
    dist.destroy_process_group()


def test_data_parallel_helpers_gloo():
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, e0, k0, m0, lg0, sg0), (r1, s1, e1, k1, m1, lg1, sg1) = res
    assert s0 == s1 and e0 == e1 and s0 != e0            # same start (broadcast), same end (averaged gradients), and they did move
    assert k0 == k1 == 3 and m0 == m1 == 1.5
    assert np.allclose(sg0, sg1) and np.allclose(sg0, (np.array(lg0) + np.array(lg1)) / 2, atol=1e-7)
    # without a process group the helpers are no-ops
    net = torch.nn.Linear(2, 2)
    net(torch.ones(1, 2)).sum().backward()
    g = net.weight.grad.clone()
    ppo.sync_gradients(list(net.parameters()))
    assert torch.equal(g, net.weight.grad) and ppo.agree_min(5) == 5 and ppo.mean_over_ranks(2.0) == 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["lf", "clifford", "pauli"])
@pytest.mark.parametrize("algo", ["PPO", "AZ"])
def test_learn_runs_for_every_env_kind(kind, algo):
    """A few iterations of each algorithm on each of the other env kinds (twists on, add_perms on for PauliNetwork): finite losses,
    sane bookkeeping, and the trained policy still synthesises through RLSynthesis.synth."""
    from qiskit_gym_b200.specs import SynthSpec
    from qiskit_gym_b200.rl import RLSynthesis

    torch.manual_seed(1)
    tri = [(0, 1), (1, 0), (1, 2), (2, 1)]
    if kind == "lf":
        env = SynthSpec.from_coupling_map("LinearFunctionEnv", tri, basis_gates=("CX",), max_depth=32)
    elif kind == "clifford":
        env = SynthSpec.from_coupling_map("CliffordEnv", tri, basis_gates=("H", "S", "CX"), max_depth=32)
    else:
        env = SynthSpec.from_coupling_map("PauliNetworkEnv", tri, basis_gates=("H", "S", "SX", "CX"), max_depth=32)
    if algo == "PPO":
        cfg = {"collecting": {"num_episodes": 128}, "training": {"num_epochs": 2}, "evals": {"ppo_deterministic": {"num_episodes": 32}}}
    else:
        cfg = {"collecting": {"num_episodes": 64, "num_mcts_searches": 6}, "training": {"num_epochs": 2}, "learning": {"diff_metric": "m"},
               "evals": {"m": {"num_episodes": 16, "num_mcts_searches": 4}}}
    rls = RLSynthesis(env, cfg, {"embedding_size": 32, "common_layers": [32]}, device=0, algorithm_cls=f"twisterl.rl.{algo}")
    hist = rls.learn(initial_difficulty=2, num_iterations=3)
    assert len(hist) == 3
    for h in hist:
        assert np.isfinite(h["loss"]) and h["samples"] > 0 and h["episodes"] > 0 and 0.0 <= h["collect_success"] <= 1.0
        assert all(0.0 <= v <= 1.0 for k, v in h.items() if k.startswith("eval/"))
