"""Scaffolding that lets the reference's UNMODIFIED Python files (qiskit_gym/envs/adapters.py, envs/synthesis.py) be imported in an
image that has neither gymnasium nor qiskit nor the Rust extension:

  * `install_third_party_stubs()` registers minimal stand-ins for the two third-party packages those files import at module
    level (only the names they touch at import time and in the code paths the tests drive; real packages win if present);
  * `oracle_rs_module()` builds a module with the four pyo3 class names whose instances are oracle envs (CPU checker), with the
    pyo3 constructor signatures of permutation.rs:266-299, linear_function.rs:373-406, clifford.rs:390-423, pauli.rs:728-775.

Test infrastructure only.
"""
from __future__ import annotations

import importlib.util
import sys
import types

import numpy as np


def _have(name):
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install_third_party_stubs():
    made = []
    if not _have("gymnasium"):
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")

        class Env:                                   # gym.Env: reset(seed=) seeds np_random; nothing else is relied on
            def reset(self, *, seed=None, options=None):
                if seed is not None:
                    self.np_random = np.random.default_rng(seed)

        class MultiBinary:
            def __init__(self, n):
                self.n = n
                self.shape = tuple(n) if hasattr(n, "__len__") else (int(n),)

        class Discrete:
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()

        gym.Env = Env
        spaces.MultiBinary, spaces.Discrete = MultiBinary, Discrete
        gym.spaces = spaces
        sys.modules["gymnasium"], sys.modules["gymnasium.spaces"] = gym, spaces
        made += ["gymnasium", "gymnasium.spaces"]
    if not _have("qiskit"):
        def mod(name, **attrs):
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
            made.append(name)
            return m

        class _Opaque:                               # placeholder types: only isinstance() checks touch them in the tested paths
            def __init__(self, *a, **k):
                raise NotImplementedError("qiskit is not installed in this image")

        class CouplingMap:
            def __init__(self, couplinglist=None):
                self._edges = [tuple(e) for e in (couplinglist or [])]

            def get_edges(self):
                return list(self._edges)

            @classmethod
            def from_line(cls, n, bidirectional=True):
                e = [(i, i + 1) for i in range(n - 1)]
                return cls(e + [(b, a) for a, b in e] if bidirectional else e)

            @classmethod
            def from_full(cls, n):
                return cls([(a, b) for a in range(n) for b in range(n) if a != b])

        class QiskitError(Exception):
            pass

        QuantumCircuit = type("QuantumCircuit", (_Opaque,), {})
        q = mod("qiskit", QuantumCircuit=QuantumCircuit)
        q.transpiler = mod("qiskit.transpiler", CouplingMap=CouplingMap)
        q.exceptions = mod("qiskit.exceptions", QiskitError=QiskitError)
        q.quantum_info = mod("qiskit.quantum_info", Clifford=type("Clifford", (_Opaque,), {}), Pauli=type("Pauli", (_Opaque,), {}))
        q.circuit = mod("qiskit.circuit")
        q.circuit.library = mod("qiskit.circuit.library")
        q.circuit.library.generalized_gates = mod("qiskit.circuit.library.generalized_gates",
                                                  LinearFunction=type("LinearFunction", (_Opaque,), {}),
                                                  PermutationGate=type("PermutationGate", (_Opaque,), {}))
    return made


def remove_stubs(made):
    for name in made:
        sys.modules.pop(name, None)


def oracle_rs_module():
    """`qiskit_gym_rs` stand-in over the CPU oracle."""
    from oracle import oracle as orc
    from qiskit_gym_b200 import _abi

    class _Base(orc.OracleEnv):
        _KIND = -1

        def reset(self, seed=0, env_id=0):           # the reference's reset() is unseeded; tests inject the Philox seed
            super().reset(seed, env_id)

    def common(kind):
        class E(_Base):
            _KIND = kind

            def __init__(self, num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=None, add_inverts=None,
                         add_perms=None, track_solution=None):
                super().__init__(kind, num_qubits, gateset, difficulty, depth_slope, max_depth, metrics_weights=metrics_weights,
                                 add_inverts=add_inverts, add_perms=add_perms, track_solution=track_solution)
        return E

    class PauliNetworkEnv(_Base):
        _KIND = _abi.ENV_PAULI_NETWORK

        def __init__(self, num_qubits, difficulty, gateset, depth_slope, max_depth, max_rotations, pauli_diff_scale=None,
                     num_qubits_decay=None, final_pauli_layers=None, metrics_weights=None, add_perms=None, pauli_layer_reward=None,
                     track_solution=None):
            super().__init__(_abi.ENV_PAULI_NETWORK, num_qubits, gateset, difficulty, depth_slope, max_depth, max_rotations=max_rotations,
                             pauli_diff_scale=pauli_diff_scale, num_qubits_decay=num_qubits_decay, final_pauli_layers=final_pauli_layers,
                             metrics_weights=metrics_weights, add_perms=add_perms, pauli_layer_reward=pauli_layer_reward,
                             track_solution=track_solution)

    m = types.ModuleType("qiskit_gym_rs_oracle")
    m.PermutationEnv = common(_abi.ENV_PERMUTATION); m.PermutationEnv.__name__ = "PermutationEnv"
    m.LinearFunctionEnv = common(_abi.ENV_LINEAR_FUNCTION); m.LinearFunctionEnv.__name__ = "LinearFunctionEnv"
    m.CliffordEnv = common(_abi.ENV_CLIFFORD); m.CliffordEnv.__name__ = "CliffordEnv"
    m.PauliNetworkEnv = PauliNetworkEnv
    return m
