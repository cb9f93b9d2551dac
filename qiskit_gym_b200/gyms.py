"""Gymnasium-facing layer of the reference, over the CUDA engine.

Mirrors src/qiskit_gym/envs/adapters.py (`gym_adapter`, 18-105) and src/qiskit_gym/envs/synthesis.py
(`BaseSynthesisEnv` 66-149, `CliffordGym` 179-217, `LinearFunctionGym` 228-259, `PermutationGym` 267-303,
`PauliGym` 372-518, `SYNTH_ENVS` 523-528): same class names, constructor signatures, `from_coupling_map`,
`from_json`, `to_json`, `get_state`, `build_circuit_from_solution`, and the Gymnasium `reset / step` contract.

Two things differ because of what this image has:
  * gymnasium is optional: with it, the wrappers are real `gym.Env`s with `MultiBinary` / `Discrete` spaces; without
    it, equivalent minimal space objects are used (same attributes `shape`, `n`, `sample`, `contains`);
  * Qiskit is optional: targets may be given as plain arrays (permutation pattern, GF(2) matrix, tableau, gate list)
    and circuits come back as gate lists `[(name, qubits)]`; when Qiskit is importable, `QuantumCircuit` / `Clifford`
    / `LinearFunction` / `PermutationGate` inputs are accepted and `QuantumCircuit`s are returned, like the reference.
"""
from __future__ import annotations

import inspect
from abc import ABC, abstractmethod
from typing import ClassVar, Iterable, List, Tuple

import numpy as np

from . import wire
from .wire import ONE_Q_GATES, ROTATION_MARKER, TWO_Q_GATES, decode_pauli_solution  # noqa: F401  (re-exported like the reference)

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    from gymnasium import spaces as _spaces
    _EnvBase = _gym.Env
    HAVE_GYMNASIUM = True
except Exception:  # gymnasium is not part of this image
    _gym = None
    HAVE_GYMNASIUM = False

    class _EnvBase:  # the two things gym.Env provides that the adapter relies on
        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.default_rng(seed)

    class _MultiBinary:
        def __init__(self, shape):
            self.shape = tuple(int(d) for d in shape)
            self.n = self.shape
            self.dtype = np.int8

        def sample(self):
            return np.random.randint(0, 2, size=self.shape).astype(np.int8)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(((x == 0) | (x == 1)).all())

    class _Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def sample(self):
            return int(np.random.randint(0, self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class _spaces:  # namespace stand-in
        MultiBinary = _MultiBinary
        Discrete = _Discrete


def gym_adapter(cls):
    """Raw env (obs_shape / observe / reward / is_final / num_actions / reset / step) -> Gymnasium env
    (adapters.py:18-105): dense int8 observation of shape obs_shape, `terminated = is_final()`, never truncated,
    stepping a final env is an assertion error, unknown attributes are forwarded to the raw env and assigning
    `difficulty` propagates to it."""

    class GymWrapper(_EnvBase):
        metadata = {"render_modes": ["human"], "render_fps": 4}

        def __init__(self, *args, **kwargs):
            self.config = kwargs.copy()
            self._raw_env = cls(*args, **kwargs)
            self._obs_shape = tuple(self._raw_env.obs_shape())
            self.observation_space = _spaces.MultiBinary(self._obs_shape)
            self.action_space = _spaces.Discrete(self._raw_env.num_actions())

        def _full_obs(self):
            full = np.zeros(int(np.prod(self._obs_shape)), dtype=np.int8)
            full[self._raw_env.observe()] = 1
            return full.reshape(self._obs_shape)

        def reset(self, *, seed=None, options=None):
            super().reset(seed=seed)
            self._raw_env.reset()
            return self._full_obs(), {}

        def step(self, action):
            assert not bool(self._raw_env.is_final()), "Action provided when env is in final state."
            self._raw_env.step(int(action))
            return self._full_obs(), float(self._raw_env.reward()), bool(self._raw_env.is_final()), False, {}

        def render(self, mode="human"):
            if hasattr(self._raw_env, "render"):
                return self._raw_env.render(mode=mode)
            if hasattr(self._raw_env, "get_state"):
                print(self._raw_env.get_state())
            else:
                print(self._full_obs())

        def close(self):
            if hasattr(self._raw_env, "close"):
                self._raw_env.close()

        def __getattr__(self, name):
            if name == "_raw_env":           # not built yet (constructor failed): avoid infinite recursion
                raise AttributeError(name)
            return getattr(self._raw_env, name)

        def __setattr__(self, name, value):
            if name in ("difficulty",) and "_raw_env" in self.__dict__:
                setattr(self._raw_env, name, value)
            else:
                super().__setattr__(name, value)

        def to_json(self):
            return self.config

    GymWrapper.__name__ = f"{cls.__name__}Gym"
    return GymWrapper


# ------------------------------------------------------------------------------------------------------
# helpers for optional Qiskit inputs
# ------------------------------------------------------------------------------------------------------
def _is_qiskit_obj(x) -> bool:
    return type(x).__module__.split(".")[0] == "qiskit"


def _is_gate_list(x) -> bool:
    return isinstance(x, (list, tuple)) and (len(x) == 0 or (isinstance(x[0], (list, tuple)) and len(x[0]) >= 2 and isinstance(x[0][0], str)))


class BaseSynthesisEnv(ABC):
    cls_name: ClassVar[str]
    allowed_gates: ClassVar[List[str]]

    @classmethod
    def from_coupling_map(cls, coupling_map, basis_gates: Tuple[str] = None, difficulty: int = 1, depth_slope: int = 2,
                          max_depth: int = 128, metrics_weights: dict | None = None, add_inverts: bool = True, add_perms: bool = True):
        """synthesis.py:71-118: edges sorted, one gate per (1-qubit gate, qubit) and per (2-qubit gate, edge) in
        basis-gate order; keyword arguments the class does not take are dropped.  `coupling_map`: a list of edges
        or anything with `get_edges()` (a Qiskit CouplingMap)."""
        if basis_gates is None:
            basis_gates = tuple(cls.allowed_gates)
        assert all(g in cls.allowed_gates for g in basis_gates), f"Some provided gates are not allowed (allowed: {cls.allowed_gates})."
        if hasattr(coupling_map, "get_edges"):
            coupling_map = list(coupling_map.get_edges())
        coupling_map = sorted(tuple(int(q) for q in e) for e in coupling_map)
        num_qubits = max(max(qubits) for qubits in coupling_map) + 1
        gateset = []
        for gate_name in basis_gates:
            if gate_name in ONE_Q_GATES:
                gateset += [(gate_name, (q,)) for q in range(num_qubits)]
            else:
                assert gate_name in TWO_Q_GATES, f"Gate {gate_name} not supported!"
                gateset += [(gate_name, (q1, q2)) for q1, q2 in coupling_map]
        config = {"num_qubits": num_qubits, "difficulty": difficulty, "gateset": gateset, "depth_slope": depth_slope,
                  "max_depth": max_depth, "metrics_weights": metrics_weights, "add_inverts": add_inverts, "add_perms": add_perms}
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        return cls(**{k: v for k, v in config.items() if k in valid})

    @classmethod
    def from_json(cls, env_config):
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        return cls(**{k: v for k, v in env_config.items() if k in valid})

    @abstractmethod
    def get_state(self, input):
        pass

    def post_process_synthesis(self, synth_circuit, _input_state):
        return synth_circuit

    def solution_gates(self, actions: List[int], input=None) -> wire.GateList:
        """Qiskit-free `build_circuit_from_solution`: the synthesised circuit as a gate list."""
        return self.post_process_synthesis(wire.solution_to_gates(self.config["gateset"], actions), input)

    def build_circuit_from_solution(self, actions: List[int], input):
        """synthesis.py:138-149.  Returns a QuantumCircuit when Qiskit is installed, else the gate list."""
        gates = self.solution_gates(actions, input)
        if wire.have_qiskit():
            return wire.gates_to_circuit(gates, self.config["num_qubits"])
        return gates


# ------------------------------------------------------------------------------------------------------
# env classes
# ------------------------------------------------------------------------------------------------------
from . import envs as _rs  # noqa: E402  (the four raw-env classes, `qiskit_gym_rs` in the reference)


def _common_kwargs(num_qubits, gateset, difficulty, depth_slope, max_depth, metrics_weights, add_inverts, add_perms, track_solution):
    return {"num_qubits": num_qubits, "difficulty": difficulty, "gateset": gateset, "depth_slope": depth_slope, "max_depth": max_depth,
            "metrics_weights": metrics_weights, "add_inverts": add_inverts, "add_perms": add_perms, "track_solution": track_solution}


# ---- Clifford ---------------------------------------------------------------------------------------
CliffordEnv = gym_adapter(_rs.CliffordEnv)


class CliffordGym(CliffordEnv, BaseSynthesisEnv):
    cls_name = "CliffordEnv"
    allowed_gates = ONE_Q_GATES + TWO_Q_GATES

    def __init__(self, num_qubits: int, gateset, difficulty: int = 1, depth_slope: int = 2, max_depth: int = 128,
                 metrics_weights: dict | None = None, add_inverts: bool = True, add_perms: bool = True, track_solution: bool = True):
        super().__init__(**_common_kwargs(num_qubits, gateset, difficulty, depth_slope, max_depth, metrics_weights, add_inverts,
                                          add_perms, track_solution))

    def target_tableau(self, input) -> np.ndarray:
        """Target as a tableau (Qiskit Clifford.tableau layout): array, gate list, or (with Qiskit) QuantumCircuit / Clifford."""
        if _is_qiskit_obj(input):
            from qiskit.quantum_info import Clifford
            return np.asarray(Clifford(input).tableau)
        if _is_gate_list(input):
            return wire.StabilizerTableau.from_gates(input, self.config["num_qubits"]).to_array()
        return np.asarray(input)

    def get_state(self, input):
        """synthesis.py:206-209: `adjoint().tableau[:, :-1].T.flatten()`."""
        return wire.clifford_state(self.target_tableau(input)).tolist()

    def post_process_synthesis(self, synth_gates, input):
        """synthesis.py:211-217: trailing Pauli layer fixing the signs (needs the target's phase column; a target given
        without phases gets no layer)."""
        if input is None:
            return synth_gates
        t = self.target_tableau(input)
        if t.shape[-1] == t.shape[-2]:
            return synth_gates
        return wire.clifford_phase_fixup(synth_gates, self.config["num_qubits"], t)


# ---- Linear function -------------------------------------------------------------------------------
LinearFunctionEnv = gym_adapter(_rs.LinearFunctionEnv)


class LinearFunctionGym(LinearFunctionEnv, BaseSynthesisEnv):
    cls_name = "LinearFunctionEnv"
    allowed_gates = ["CX", "SWAP"]

    def __init__(self, num_qubits: int, gateset, difficulty: int = 1, depth_slope: int = 2, max_depth: int = 128,
                 metrics_weights: dict | None = None, add_inverts: bool = True, add_perms: bool = True, track_solution: bool = True):
        super().__init__(**_common_kwargs(num_qubits, gateset, difficulty, depth_slope, max_depth, metrics_weights, add_inverts,
                                          add_perms, track_solution))

    def target_matrix(self, input) -> np.ndarray:
        if _is_gate_list(input):
            n = self.config["num_qubits"]
            M = np.eye(n, dtype=np.uint8)
            for name, qs in input:
                g = name.lower()
                if g == "cx":
                    M[qs[1]] ^= M[qs[0]]
                elif g == "swap":
                    M[[qs[0], qs[1]]] = M[[qs[1], qs[0]]]
                else:
                    raise TypeError(f"Gate {name} on qubits {list(qs)} not supported.")
            return M
        return np.asarray(input)

    def get_state(self, input):
        """synthesis.py:255-259: the linear matrix of the inverse circuit.  input: {0,1}[n, n] matrix of the linear function,
        a CX/SWAP gate list, or (with Qiskit) a QuantumCircuit / LinearFunction."""
        if _is_qiskit_obj(input):
            from qiskit.circuit.library.generalized_gates import LinearFunction
            from qiskit.quantum_info import Clifford
            return np.array(LinearFunction(Clifford(input).adjoint()).linear).flatten().astype(int).tolist()
        return wire.linear_function_state(self.target_matrix(input)).tolist()


# ---- Permutation -----------------------------------------------------------------------------------
PermutationEnv = gym_adapter(_rs.PermutationEnv)


class PermutationGym(PermutationEnv, BaseSynthesisEnv):
    cls_name = "PermutationEnv"
    allowed_gates = ["SWAP"]

    def __init__(self, num_qubits: int, gateset, difficulty: int = 1, depth_slope: int = 2, max_depth: int = 128,
                 metrics_weights: dict | None = None, add_inverts: bool = True, add_perms: bool = True, track_solution: bool = True):
        super().__init__(**_common_kwargs(num_qubits, gateset, difficulty, depth_slope, max_depth, metrics_weights, add_inverts,
                                          add_perms, track_solution))

    def get_state(self, input):
        """synthesis.py:294-303: argsort of the pattern (the inverse permutation)."""
        if _is_qiskit_obj(input):
            from qiskit import QuantumCircuit
            from qiskit.circuit.library.generalized_gates import LinearFunction
            input = LinearFunction(input).permutation_pattern() if isinstance(input, QuantumCircuit) else input.pattern
        return np.argsort(np.array(input)).astype(int).tolist()


# ---- Pauli network ---------------------------------------------------------------------------------
PauliNetworkEnv = gym_adapter(_rs.PauliNetworkEnv)


class PauliGym(PauliNetworkEnv, BaseSynthesisEnv):
    cls_name = "PauliNetworkEnv"
    allowed_gates = ONE_Q_GATES + TWO_Q_GATES

    def __init__(self, num_qubits: int, gateset, difficulty: int = 1, depth_slope: int = 2, max_depth: int = 128, max_rotations: int = 5,
                 pauli_diff_scale: int = 16, num_qubits_decay: float = 0.5, final_pauli_layers: int | None = None,
                 metrics_weights: dict | None = None, add_perms: bool = True, pauli_layer_reward: float = 0.01, track_solution: bool = True):
        super().__init__(**{
            "num_qubits": num_qubits, "difficulty": difficulty, "gateset": gateset, "depth_slope": depth_slope, "max_depth": max_depth,
            "max_rotations": max_rotations, "pauli_diff_scale": pauli_diff_scale, "num_qubits_decay": num_qubits_decay,
            "final_pauli_layers": final_pauli_layers, "metrics_weights": metrics_weights, "add_perms": add_perms,
            "pauli_layer_reward": pauli_layer_reward, "track_solution": track_solution})
        self._rotation_params = []
        self._original_circuit = None

    def get_state(self, input, rotations: List[str] = None):
        """synthesis.py:414-459.  input: (tableau, rotation labels) tuple — tableau already the adjoint, as in the reference —
        or a bare tableau array with `rotations` (adjoint taken here, like the reference's raw-Clifford branch).  QuantumCircuit /
        Clifford inputs need Qiskit (circuit parsing evolves Paulis through Cliffords, synthesis.py:320-361)."""
        if _is_qiskit_obj(input) or (isinstance(input, tuple) and len(input) == 2 and _is_qiskit_obj(input[0])):
            from ._qiskit_bridge import pauli_get_state
            return pauli_get_state(self, input, rotations)
        self._rotation_params = []
        self._original_circuit = None
        if isinstance(input, tuple):
            tableau, rotations = input
            return wire.pauli_network_state(tableau, list(rotations), adjoint=False)
        if isinstance(input, np.ndarray):
            return wire.pauli_network_state(input, list(rotations or []), adjoint=True)
        raise ValueError(f"Unsupported input type: {type(input)}")

    def set_rotation_params(self, params):
        """Angles of the target's rotations, in label order (the reference fills these while parsing a QuantumCircuit)."""
        self._rotation_params = list(params)

    def solution_gates(self, actions, input=None):
        return wire.pauli_solution_to_gates(self.config["gateset"], actions, self._rotation_params or None)

    def build_circuit_from_solution(self, actions, input):
        """synthesis.py:502-518."""
        if wire.have_qiskit():
            from ._qiskit_bridge import pauli_reconstruct
            return pauli_reconstruct(self, decode_pauli_solution(actions), input)
        return self.solution_gates(actions, input)


SYNTH_ENVS = {
    "CliffordEnv": CliffordGym,
    "LinearFunctionEnv": LinearFunctionGym,
    "PermutationEnv": PermutationGym,
    "PauliNetworkEnv": PauliGym,
}
