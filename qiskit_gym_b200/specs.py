"""`SynthSpec`: what `RLSynthesis` needs to know about a synthesis problem, without Gymnasium or Qiskit.

In the reference that role is played by the `*Gym` objects of src/qiskit_gym/envs/synthesis.py (a Gymnasium wrapper around one
Rust env, plus `get_state` / `build_circuit_from_solution` on Qiskit objects).  Those classes are NOT restated here: with the
reference's Python package present they are used as they are (reference_shim.install() binds them to the engine).  A
`SynthSpec` is the Qiskit-free stand-in for machines without it: a config dict in the reference's JSON schema
(`RLSynthesis.save`, rl/synthesis.py:79-93) plus the wire formats of wire.py —

    to_json / cls_name / obs_shape / num_actions        what RLSynthesis reads from an env object (rl/synthesis.py:43, 97-101)
    get_state(target)                                   plain arrays / gate lists -> the Vec<i64> payload of Env::set_state
    build_circuit_from_solution(actions, target)        Env::solution -> gate list [(name, qubits)]

It owns no environment: searches and collectors build their own BatchedEnv from the config.
"""
from __future__ import annotations

import numpy as np

from . import _abi, host, wire
from . import workloads as W

_KINDS = {"PermutationEnv": _abi.ENV_PERMUTATION, "LinearFunctionEnv": _abi.ENV_LINEAR_FUNCTION, "CliffordEnv": _abi.ENV_CLIFFORD,
          "PauliNetworkEnv": _abi.ENV_PAULI_NETWORK}
_ALLOWED = {"PermutationEnv": ("SWAP",), "LinearFunctionEnv": ("CX", "SWAP"), "CliffordEnv": W.ALL_GATES, "PauliNetworkEnv": W.ALL_GATES}
_ENGINE_KEYS = ("metrics_weights", "add_inverts", "add_perms", "track_solution", "max_rotations", "pauli_diff_scale", "num_qubits_decay",
                "final_pauli_layers", "pauli_layer_reward")


def _is_gate_list(x) -> bool:
    return isinstance(x, (list, tuple)) and (len(x) == 0 or (isinstance(x[0], (list, tuple)) and len(x[0]) >= 2 and isinstance(x[0][0], str)))


class SynthSpec:
    def __init__(self, cls_name: str, config: dict):
        if cls_name not in _KINDS:
            raise ValueError(f"Synth env class {cls_name} not supported, should be {list(_KINDS)}")
        self.cls_name = cls_name
        self.kind = _KINDS[cls_name]
        self.config = dict(config)                       # kept exactly as given: to_json() round-trips the reference's JSON files
        self.gateset = [(g, tuple(int(q) for q in qs)) for g, qs in config["gateset"]]
        kw = {k: self.config[k] for k in _ENGINE_KEYS if k in self.config}
        if self.kind != _abi.ENV_PAULI_NETWORK:
            kw = {k: v for k, v in kw.items() if k in ("metrics_weights", "add_inverts", "add_perms", "track_solution")}
        else:
            kw.pop("add_inverts", None)
            kw.setdefault("max_rotations", 5)
        self._cfg = host.make_config(self.kind, self.config["num_qubits"], self.gateset, self.config.get("difficulty", 1),
                                     self.config.get("depth_slope", 2), self.config.get("max_depth", 128), **kw)
        host.validate(self._cfg)
        self._rotation_params = []

    # ---- construction ------------------------------------------------------------------------------------------------
    @classmethod
    def from_coupling_map(cls, cls_name: str, coupling_map, basis_gates=None, **config):
        """Gateset in the reference's order (envs/synthesis.py:91-103, pinned by tests/golden/reference_python_kats.json);
        remaining keyword arguments go into the config dict unchanged."""
        basis = tuple(basis_gates) if basis_gates is not None else _ALLOWED[cls_name]
        bad = [g for g in basis if g not in _ALLOWED[cls_name]]
        if bad:
            raise ValueError(f"gates {bad} are not allowed for {cls_name} (allowed: {list(_ALLOWED[cls_name])})")
        edges = coupling_map.get_edges() if hasattr(coupling_map, "get_edges") else coupling_map
        n, gateset = W.gateset_from_coupling_map([tuple(int(q) for q in e) for e in edges], basis)
        full = {"num_qubits": n, "difficulty": 1, "gateset": gateset, "depth_slope": 2, "max_depth": 128}
        full.update(config)
        return cls(cls_name, full)

    @classmethod
    def from_json(cls, cls_name: str, env_config: dict):
        return cls(cls_name.split(".")[-1], env_config)

    # ---- what RLSynthesis reads ----------------------------------------------------------------------------------------
    def to_json(self) -> dict:
        return self.config

    def obs_shape(self):
        return host.obs_shape(self._cfg)

    def num_actions(self) -> int:
        return len(self.gateset)

    # ---- wire formats -----------------------------------------------------------------------------------------------------
    def _tableau(self, target) -> np.ndarray:
        if _is_gate_list(target):
            return wire.StabilizerTableau.from_gates(target, self.config["num_qubits"]).to_array()
        return np.asarray(target)

    def get_state(self, target, rotations=None):
        """Permutation: pattern int[n].  LinearFunction: {0,1}[n, n] matrix or a CX/SWAP gate list.  Clifford: tableau
        bool[2n, 2n(+1)] (Qiskit's Clifford.tableau layout) or a gate list.  PauliNetwork: (tableau, labels) with the tableau already
        the adjoint, or a tableau with `rotations=` (adjoint taken here) — the two array branches of envs/synthesis.py:414-459."""
        n = self.config["num_qubits"]
        if self.kind == _abi.ENV_PERMUTATION:
            return wire.permutation_state(target).tolist()
        if self.kind == _abi.ENV_LINEAR_FUNCTION:
            if _is_gate_list(target):
                M = np.eye(n, dtype=np.uint8)
                for name, qs in target:
                    g = name.lower()
                    if g == "cx":
                        M[qs[1]] ^= M[qs[0]]
                    elif g == "swap":
                        M[[qs[0], qs[1]]] = M[[qs[1], qs[0]]]
                    else:
                        raise TypeError(f"Gate {name} on qubits {list(qs)} not supported.")
                target = M
            return wire.linear_function_state(np.asarray(target)).tolist()
        if self.kind == _abi.ENV_CLIFFORD:
            return wire.clifford_state(self._tableau(target)).tolist()
        self._rotation_params = []
        if isinstance(target, tuple):
            tableau, rotations = target
            return wire.pauli_network_state(self._tableau(tableau), list(rotations), adjoint=False)
        return wire.pauli_network_state(self._tableau(target), list(rotations or []), adjoint=True)

    def set_rotation_params(self, params):
        """Angles of a PauliNetwork target's rotations, in label order."""
        self._rotation_params = list(params)

    def build_circuit_from_solution(self, actions, target=None):
        """Env::solution -> gate list.  Clifford targets that carry a phase column get the trailing Pauli layer that fixes the signs
        (what envs/synthesis.py:162-177, 211-217 does with Qiskit objects); PauliNetwork solutions are decoded into gates and rotations."""
        gs = self.gateset
        if self.kind == _abi.ENV_PAULI_NETWORK:
            return wire.pauli_solution_to_gates(gs, actions, self._rotation_params or None)
        gates = wire.solution_to_gates(gs, actions)
        if self.kind == _abi.ENV_CLIFFORD and target is not None:
            t = self._tableau(target)
            if t.shape[-1] != t.shape[-2]:
                return wire.clifford_phase_fixup(gates, self.config["num_qubits"], t)
        return gates
