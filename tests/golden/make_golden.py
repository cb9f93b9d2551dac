#!/usr/bin/env python
"""Regenerates tests/golden/notebook_kats.json from the reference's own tutorial notebook
(/root/reference/examples/intro.ipynb) — the only executable record of reference behaviour in the tree.

Run in the build container (the notebook is not available on the GPU box):
    python tests/golden/make_golden.py
Extracted: gateset orderings printed by from_coupling_map (cells 3, 16), the state rendered after
set_state (cell 7), observations / final flags printed by env.step (cells 10-12), the space shapes
(cells 8-9) and the difficulty-1 reset observation (cell 4).  Rewards printed there are from an older reward
scheme and are recorded as `stale_reward` only.
Also records the parameter shapes of the three saved BasicPolicy checkpoints (examples/models/*.pt).
"""
import ast
import json
import os
import re

NB = "/root/reference/examples/intro.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_kats.json")


def cell_output(cell):
    txt = ""
    for o in cell.get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain") or ""
        txt += "".join(t)
    return txt


def matrices(txt):
    """all [[..] [..]] integer blocks in printed order"""
    out = []
    for m in re.finditer(r"\[\[[0-9\s\[\],]+?\]\]", txt):
        rows = re.findall(r"\[([0-9\s,]+)\]", m.group(0))
        out.append([[int(x) for x in re.split(r"[\s,]+", r.strip()) if x] for r in rows])
    return out


def main():
    nb = json.load(open(NB))
    cells = nb["cells"]
    src = lambda i: "".join(cells[i]["source"])
    k = {"source": "examples/intro.ipynb"}
    # cell 3: LF gateset on a bidirectional 3-line
    assert "from_line(3" in src(3)
    k["lf3_gateset"] = [[g, list(q)] for g, q in ast.literal_eval(cell_output(cells[3]))]
    # cell 4: reset at difficulty 1
    k["lf3_reset_difficulty1_obs"] = matrices(cell_output(cells[4]))[0]
    # cell 7: render after set_state(get_state(cx(0,2)))
    k["lf3_state_after_set_state"] = matrices(cell_output(cells[7]))[0]
    k["lf3_action_space"] = int(re.search(r"Discrete\((\d+)\)", cell_output(cells[8])).group(1))
    k["lf3_obs_space"] = [int(x) for x in re.search(r"MultiBinary\(\((\d+), (\d+)\)\)", cell_output(cells[9])).groups()]
    # cell 10: step(2)
    out10 = cell_output(cells[10])
    k["lf3_step2"] = {"action": 2, "obs": matrices(out10)[0], "is_final": "True" in out10.split("\n")[-1],
                      "stale_reward": float(re.search(r"(-?\d+\.\d+),", out10.split("dtype=int8),")[1]).group(1))}
    # cells 11, 12: sequences
    for ci, key in ((11, "lf3_sequence_a"), (12, "lf3_sequence_b")):
        out = cell_output(cells[ci])
        acts = [int(a) for a in re.findall(r"^\[(\d+)\] - \(", out, flags=re.M)]
        finals = [s == "True" for s in re.findall(r"Is final: (True|False)", out)]
        mats = matrices(out)
        k[key] = {"start": mats[0], "actions": acts, "states": mats[1:], "is_final": finals}
    # cell 16: 3x3 grid permutation gateset
    assert "from_grid(3,3" in src(16)
    k["perm_grid3_gateset"] = [[g, list(q)] for g, q in ast.literal_eval(cell_output(cells[16]))]
    # checkpoints: parameter shapes only
    try:
        import torch
        shapes = {}
        for name in ("perm_square_3x3", "lf_5_line", "clifford_3q_custom"):
            sd = torch.load(f"/root/reference/examples/models/{name}.pt", weights_only=True, map_location="cpu")
            shapes[name] = {p: list(v.shape) for p, v in sd.items()}
        k["policy_checkpoint_shapes"] = shapes
    except Exception as ex:  # pragma: no cover
        k["policy_checkpoint_shapes"] = {"error": str(ex)}
    json.dump(k, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()


def copy_model_fixtures():
    """The reference's three trained BasicPolicy checkpoints + RLSynthesis config files (examples/models/*.pt, *.json) are data
    fixtures of the synth path: tests/test_gyms.py loads them through RLSynthesis.from_config_json and checks that the
    device-resident search synthesises valid circuits with them.  Copied verbatim (they are model data, not source)."""
    import shutil
    src = "/root/reference/examples/models"
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(src)):
        if f.endswith((".pt", ".json")):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))


if __name__ == "__main__" and os.path.isdir("/root/reference/examples/models"):
    copy_model_fixtures()


def make_config_kats(out_path):
    """config_kats.json: `to_json()` outputs of the reference's own config classes (src/qiskit_gym/rl/configs.py has no third-party
    dependency, so it is imported as a plain module) for defaults, custom values and a partial `from_json`."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_configs", "/root/reference/src/qiskit_gym/rl/configs.py")
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_configs"] = m
    spec.loader.exec_module(m)
    out = {
        "ppo_default": m.PPOConfig().to_json(), "az_default": m.AlphaZeroConfig().to_json(),
        "basic_default": m.BasicPolicyConfig().to_json(), "conv_default": m.Conv1dPolicyConfig().to_json(),
        "ppo_custom": m.PPOConfig(num_episodes=64, gae_lambda=0.9, lr=1e-3, diff_max=12, evals={"quick": m.EvalConfig(num_episodes=8)}, diff_metric="quick").to_json(),
        "az_custom": m.AlphaZeroConfig(num_mcts_searches=32, C=2.0, evals={"m": m.EvalConfig(num_mcts_searches=8)}, diff_metric="m").to_json(),
        "ppo_from_partial": m.PPOConfig.from_json({"collecting": {"num_episodes": 7}, "evals": {"x": {"num_searches": 3}}}).to_json(),
        "basic_custom": m.BasicPolicyConfig(embedding_size=64, common_layers=[32, 16], value_layers=[8]).to_json(),
    }
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__" and os.path.isfile("/root/reference/src/qiskit_gym/rl/configs.py"):
    import sys
    make_config_kats(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_kats.json"))
