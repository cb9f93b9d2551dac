#!/usr/bin/env python
"""Generates tests/golden/reference_python_kats.json by EXECUTING the reference's own Python layer, unmodified, from
/root/reference/src (qiskit_gym/envs/adapters.py and qiskit_gym/envs/synthesis.py), with
  * `qiskit_gym.qiskit_gym_rs` bound to the CPU oracle's env classes (tests/ref_stubs.py) through
    qiskit_gym_b200.reference_shim.install — the same mechanism that binds it to the CUDA engine in production,
  * stand-ins for gymnasium / qiskit (neither is installed here; tests/ref_stubs.py).

What the reference code itself computes and this file records:
  - `XGym.from_coupling_map` (envs/synthesis.py:71-118): gateset order, filtered constructor config, `to_json()`;
  - `gym_adapter` (adapters.py:18-105): observation / action spaces, the dense int8 observation built from `observe()`, the
    `(obs, reward, terminated, truncated, info)` tuples of `step`, `reset`, `difficulty` forwarding, the assertion on stepping a
    final env — recorded as whole-episode traces;
  - `PermutationGym.get_state` on plain patterns (envs/synthesis.py:294-303);
  - `decode_pauli_solution` (envs/synthesis.py:35-61) on the encoded solutions of PauliNetwork episodes.
The traces pin the raw-env protocol the wrappers rely on; tests/test_reference_python.py replays them on the CUDA engine.

    python tests/golden/make_reference_python_golden.py [--check]
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_SRC = os.environ.get("QISKIT_GYM_SRC", "/root/reference/src")
OUT = os.path.join(HERE, "reference_python_kats.json")

# (gym class, edges, basis gates, extra from_coupling_map kwargs, target scramble, actions to play, seed)
LINE3 = [(0, 1), (1, 0), (1, 2), (2, 1)]
LINE4 = [(0, 1), (1, 2), (2, 3)]
GRID3 = [(0, 1), (0, 3), (1, 2), (1, 4), (2, 5), (3, 4), (3, 6), (4, 5), (4, 7), (5, 8), (6, 7), (7, 8)]
TRI = [(0, 1), (1, 0), (1, 2), (2, 1), (0, 2), (2, 0)]
CASES = [
    ("PermutationGym", GRID3, None, dict(max_depth=14), 6, 14, 11),
    ("PermutationGym", LINE4, None, dict(max_depth=8, depth_slope=3), 3, 8, 12),
    ("LinearFunctionGym", LINE3, ("CX",), dict(max_depth=10), 4, 10, 21),
    ("LinearFunctionGym", [(3, 2), (0, 1), (2, 1), (1, 0), (1, 2), (2, 3)], ("SWAP", "CX"), dict(max_depth=12, metrics_weights={"n_cnots": 0.02, "n_layers": 0.005}), 5, 12, 22),
    ("CliffordGym", TRI, ("H", "S", "CX"), dict(max_depth=12), 5, 12, 31),
    ("CliffordGym", LINE3, None, dict(max_depth=16, metrics_weights={"n_gates": 0.001, "n_layers_cnots": 0.01}), 7, 16, 32),
    ("PauliGym", TRI, ("H", "S", "SX", "CX"), dict(max_depth=20), 6, 20, 41),
    ("PauliGym", LINE4 + [(b, a) for a, b in LINE4], None, dict(max_depth=24), 8, 24, 42),
]


def load_reference(backend):
    from qiskit_gym_b200 import reference_shim as shim
    shim.uninstall()
    shim.install(REF_SRC, backend=backend)
    import qiskit_gym.envs.synthesis as syn           # the reference's file
    assert os.path.realpath(syn.__file__).startswith(os.path.realpath(REF_SRC)), syn.__file__
    return syn


def f32_bits(x) -> int:
    return int(np.float32(x).view(np.uint32))


def dense_to_idx(obs) -> list:
    return np.flatnonzero(np.asarray(obs).reshape(-1)).tolist()


def case_inputs(syn, cls_name, edges, basis, kw, scramble, steps, seed):
    """Deterministic target payload and action list of a case (also used by the GPU test)."""
    from qiskit_gym_b200 import workloads as W
    kind = {"PermutationGym": W.PERM, "LinearFunctionGym": W.LF, "CliffordGym": W.CLIFF, "PauliGym": W.PAULI}[cls_name]
    allowed = {"PermutationGym": ("SWAP",), "LinearFunctionGym": ("CX", "SWAP"), "CliffordGym": W.ALL_GATES, "PauliGym": W.ALL_GATES}[cls_name]
    n, gateset = W.gateset_from_coupling_map(edges, basis or allowed)
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == W.PAULI:
        state = W.random_targets(kind, n, gateset, 1, seed, scramble=scramble, num_rotations=3)[0]
        state = [int(v) for v in state]
        actions = rng.integers(0, len(gateset), size=steps).tolist()
    else:
        # identity scrambled by `scramble` gates, then the same gates backwards: every gate is an involution on the GF(2) state
        # (permutation.rs:205-208, linear_function.rs:62-83, clifford.rs:89-133), so the episode ends in success after `scramble` steps
        fwd = rng.integers(0, len(gateset), size=scramble).tolist()
        actions = fwd[::-1] + rng.integers(0, len(gateset), size=max(steps - scramble, 0)).tolist()
        rs = syn.qiskit_gym_rs
        raw_cls = {W.PERM: rs.PermutationEnv, W.LF: rs.LinearFunctionEnv, W.CLIFF: rs.CliffordEnv}[kind]
        probe = raw_cls(n, 1, gateset, 2, 10 ** 6, None, False, False, True)
        for a in fwd:
            probe.step(a)
        raw = probe.raw_state()
        state = [int(v) for v in raw]
    return n, gateset, state, [int(a) for a in actions]


def run_case(syn, case):
    cls_name, edges, basis, kw, scramble, steps, seed = case
    cls = getattr(syn, cls_name)
    fk = dict(kw)
    if cls_name != "PauliGym":
        fk["add_inverts"] = False
    fk["add_perms"] = False
    env = cls.from_coupling_map(edges, basis_gates=basis, **fk)
    n, gateset, state, actions = case_inputs(syn, cls_name, edges, basis, kw, scramble, steps, seed)
    rec = {
        "cls": cls_name, "edges": [list(e) for e in edges], "basis_gates": None if basis is None else list(basis), "kwargs": fk,
        "scramble": scramble, "steps": steps, "seed": seed,
        "wrapper_name": type(env).__mro__[1].__name__, "cls_name": env.cls_name,
        "config": json.loads(json.dumps(env.to_json())),
        "observation_space_shape": list(env.observation_space.shape), "action_space_n": int(env.action_space.n),
        "state": state, "actions": actions,
    }
    assert [tuple(g[1]) for g in env.config["gateset"]] == [tuple(g[1]) for g in gateset]
    env.set_state(state)                                   # forwarded to the raw env by GymWrapper.__getattr__
    rec["obs0"] = dense_to_idx(env._full_obs())
    rec["obs_dtype"] = str(env._full_obs().dtype)
    trace = []
    for a in actions:
        if env.is_final():
            try:
                env.step(a)
                raise RuntimeError("the reference wrapper must assert on a final env")
            except AssertionError as ex:
                rec["final_assert"] = str(ex)
            break
        obs, reward, terminated, truncated, info = env.step(a)
        assert truncated is False and info == {}
        trace.append({"a": a, "obs": dense_to_idx(obs), "reward_bits": f32_bits(reward), "terminated": bool(terminated)})
    rec["trace"] = trace
    rec["success"] = bool(env.success())
    rec["solution"] = [int(v) for v in env.solution()]
    if cls_name == "PauliGym":
        rec["decoded_solution"] = [list(t) for t in syn.decode_pauli_solution(rec["solution"])]
    # difficulty forwarding + reset contract
    env.difficulty = 3
    rec["difficulty_after_set"] = int(env._raw_env.difficulty)
    obs, info = env.reset(seed=5)
    rec["reset_returns"] = [type(obs).__name__, list(obs.shape), info]
    return rec


def generate():
    from tests import ref_stubs
    made = ref_stubs.install_third_party_stubs()
    try:
        syn = load_reference(ref_stubs.oracle_rs_module())
        out = {
            "generator": "tests/golden/make_reference_python_golden.py",
            "reference_files": ["src/qiskit_gym/envs/adapters.py", "src/qiskit_gym/envs/synthesis.py"],
            "backend": "oracle (CPU restatement) bound as qiskit_gym.qiskit_gym_rs",
            "constants": {"ONE_Q_GATES": syn.ONE_Q_GATES, "TWO_Q_GATES": syn.TWO_Q_GATES, "ROTATION_MARKER": syn.ROTATION_MARKER,
                          "SYNTH_ENVS": {k: v.__name__ for k, v in syn.SYNTH_ENVS.items()},
                          "allowed_gates": {v.__name__: list(v.allowed_gates) for v in syn.SYNTH_ENVS.values()}},
            "cases": [run_case(syn, c) for c in CASES],
        }
        # PermutationGym.get_state on plain patterns (envs/synthesis.py:294-303)
        penv = syn.PermutationGym.from_coupling_map(LINE4, add_inverts=False, add_perms=False)
        rng = np.random.Generator(np.random.PCG64(7))
        out["perm_get_state"] = [{"pattern": p, "state": penv.get_state(p)} for p in (rng.permutation(4).tolist() for _ in range(6))]
        # decode_pauli_solution on hand-made words covering every field (envs/synthesis.py:35-61)
        words = [0, 5, 0x7FFFFFFF, 0x80000000, 0x80000001, 0x80000000 | (1 << 21) | (9 << 11) | (3 << 1) | 1, 0x80000000 | (2 << 21) | (1023 << 11) | (1023 << 1)]
        out["decode_words"] = {"words": words, "decoded": [list(t) for t in syn.decode_pauli_solution(words)]}
        # from_json round trip and the kwargs filter (envs/synthesis.py:120-128)
        cfg = dict(out["cases"][0]["config"]); cfg["unknown_key"] = 1
        again = syn.PermutationGym.from_json(cfg)
        out["from_json_config"] = json.loads(json.dumps(again.to_json()))
        return out
    finally:
        from qiskit_gym_b200 import reference_shim as shim
        shim.uninstall()
        ref_stubs.remove_stubs(made)


def main():
    data = generate()
    text = json.dumps(data, indent=None, separators=(",", ":"), sort_keys=True)
    if "--check" in sys.argv:
        assert json.loads(open(OUT).read()) == json.loads(text), "fixture differs from the live reference run"
        print("fixture matches the live reference run")
        return
    with open(OUT, "w") as f:
        f.write(text + "\n")
    print("wrote", OUT, len(text), "bytes;", len(data["cases"]), "cases")


if __name__ == "__main__":
    main()
