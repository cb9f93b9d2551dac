"""ctypes mirror of include/qg_engine.h (structs, enums) and the gateset/config marshalling
shared by the engine binding.  Nothing here computes anything: it only lays out arguments."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

ENV_PERMUTATION, ENV_LINEAR_FUNCTION, ENV_CLIFFORD, ENV_PAULI_NETWORK = 0, 1, 2, 3
ENV_NAMES = {
    ENV_PERMUTATION: "PermutationEnv",
    ENV_LINEAR_FUNCTION: "LinearFunctionEnv",
    ENV_CLIFFORD: "CliffordEnv",
    ENV_PAULI_NETWORK: "PauliNetworkEnv",
}
GATE_NAMES = ["H", "S", "Sdg", "SX", "SXdg", "CX", "CZ", "SWAP"]

QG_OK, QG_ERR_INVALID, QG_ERR_CUDA, QG_ERR_UNSUPPORTED, QG_ERR_STATE = 0, -1, -2, -3, -4

FLAG_SINGULAR, FLAG_SOLUTION_OVERFLOW, FLAG_LAYER_OVERFLOW, FLAG_BAD_ROTATION, FLAG_BAD_ACTION = 1, 2, 4, 8, 16


class QgGate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("q0", C.c_int32), ("q1", C.c_int32)]


class QgConfig(C.Structure):
    _fields_ = [
        ("env_kind", C.c_int32),
        ("num_qubits", C.c_int32),
        ("difficulty", C.c_int32),
        ("depth_slope", C.c_int32),
        ("max_depth", C.c_int32),
        ("num_gates", C.c_int32),
        ("gateset", C.POINTER(QgGate)),
        ("w_n_cnots", C.c_float),
        ("w_n_layers_cnots", C.c_float),
        ("w_n_layers", C.c_float),
        ("w_n_gates", C.c_float),
        ("add_inverts", C.c_int32),
        ("add_perms", C.c_int32),
        ("track_solution", C.c_int32),
        ("max_rotations", C.c_int32),
        ("pauli_diff_scale", C.c_int32),
        ("num_qubits_decay", C.c_float),
        ("final_pauli_layers", C.c_int32),
        ("pauli_layer_reward", C.c_float),
        ("solution_capacity", C.c_int32),
        ("tile_envs", C.c_int32),
    ]


class QgMctsTree(C.Structure):
    """qg_mcts_tree: device arrays of the PUCT trees (include/qg_engine.h)."""
    _fields_ = [
        ("num_trees", C.c_int32), ("node_cap", C.c_int32), ("num_actions", C.c_int32),
        ("prior", C.c_void_p), ("visits", C.c_void_p), ("value_sum", C.c_void_p), ("child", C.c_void_p),
        ("node_reward", C.c_void_p), ("node_final", C.c_void_p), ("node_count", C.c_void_p),
        ("path_node", C.c_void_p), ("path_action", C.c_void_p), ("path_len", C.c_void_p), ("new_node", C.c_void_p),
    ]


def parse_gateset(gateset: Iterable, kind_from_name) -> "C.Array[QgGate]":
    """(name, indices) pairs -> qg_gate[]; error behaviour of common.rs:46-100
    (TypeError for malformed items, ValueError for unknown names / wrong arity)."""
    items = list(gateset)
    arr = (QgGate * max(len(items), 1))()
    for i, item in enumerate(items):
        try:
            pair = list(item)
        except TypeError:
            raise TypeError("Each gate must be a 2-item sequence: (name, indices)")
        if isinstance(item, (str, bytes)):
            raise TypeError("Each gate must be a 2-item sequence: (name, indices)")
        if len(pair) != 2:
            raise ValueError("Each gate must have exactly 2 items: (name, indices)")
        name, idx = pair
        if not isinstance(name, str):
            raise TypeError("Gate name must be a string")
        if isinstance(idx, (str, bytes)):
            raise TypeError("Gate indices must be a list/tuple of integers")
        try:
            idx = list(idx)
        except TypeError:
            raise TypeError("Gate indices must be a list/tuple of integers")
        qs = []
        for q in idx:
            if isinstance(q, bool) or not isinstance(q, int) and not hasattr(q, "__index__"):
                raise TypeError("Gate indices must be non-negative integers (usize)")
            q = int(q)
            if q < 0:
                raise TypeError("Gate indices must be non-negative integers (usize)")
            qs.append(q)
        kind = kind_from_name(name.encode(), len(qs))
        if kind == QG_ERR_INVALID:
            raise ValueError(
                f"Unknown gate name `{name.strip()}`. Allowed: H, S, Sdg, SX, SXdg, CX, CZ, SWAP"
            )
        if kind == QG_ERR_STATE:
            nm = name.strip()
            want = 2 if nm.lower() in ("cx", "cz", "swap") else 1
            raise ValueError(
                f"Gate `{nm}` expects {want} ind{'ices' if want == 2 else 'ex'}, got {len(qs)}"
            )
        arr[i].kind = kind
        arr[i].q0 = qs[0]
        arr[i].q1 = qs[1] if len(qs) > 1 else 0
    return arr


def make_config(
    env_kind: int,
    num_qubits: int,
    difficulty: int,
    gates: "C.Array[QgGate]",
    num_gates: int,
    depth_slope: int,
    max_depth: int,
    metrics_weights: dict | None = None,
    add_inverts: bool | None = None,
    add_perms: bool | None = None,
    track_solution: bool | None = None,
    max_rotations: int = 5,
    pauli_diff_scale: int | None = None,
    num_qubits_decay: float | None = None,
    final_pauli_layers: int | None = None,
    pauli_layer_reward: float | None = None,
    solution_capacity: int = 0,
    tile_envs: int = 0,
) -> QgConfig:
    """Resolves the pyo3 `None` defaults exactly like the reference constructors
    (permutation.rs:277-299, pauli.rs:743-775; MetricsWeights::from_hashmap metrics.rs:169-184)."""
    cfg = QgConfig()
    cfg.env_kind = env_kind
    cfg.num_qubits = int(num_qubits)
    cfg.difficulty = int(difficulty)
    cfg.depth_slope = int(depth_slope)
    cfg.max_depth = int(max_depth)
    cfg.num_gates = int(num_gates)
    cfg.gateset = C.cast(gates, C.POINTER(QgGate))
    w = {"n_cnots": 0.01, "n_layers_cnots": 0.0, "n_layers": 0.0, "n_gates": 0.0001}
    if metrics_weights:
        for k, v in metrics_weights.items():
            if k in w:
                w[k] = float(v)
    cfg.w_n_cnots, cfg.w_n_layers_cnots = w["n_cnots"], w["n_layers_cnots"]
    cfg.w_n_layers, cfg.w_n_gates = w["n_layers"], w["n_gates"]
    cfg.add_inverts = 1 if (True if add_inverts is None else add_inverts) else 0
    if env_kind == ENV_PAULI_NETWORK:
        cfg.add_inverts = 0
    cfg.add_perms = 1 if (True if add_perms is None else add_perms) else 0
    cfg.track_solution = 1 if (True if track_solution is None else track_solution) else 0
    cfg.max_rotations = int(max_rotations)
    cfg.pauli_diff_scale = 8 if pauli_diff_scale is None else int(pauli_diff_scale)
    cfg.num_qubits_decay = 0.5 if num_qubits_decay is None else float(num_qubits_decay)
    cfg.final_pauli_layers = (int(max_rotations) + 2) if final_pauli_layers is None else int(final_pauli_layers)
    cfg.pauli_layer_reward = 0.01 if pauli_layer_reward is None else float(pauli_layer_reward)
    cfg.solution_capacity = int(solution_capacity)
    cfg.tile_envs = int(tile_envs)
    cfg._keepalive = gates  # the struct only borrows the pointer
    return cfg


def as_i64_array(values: Sequence[int]):
    arr = (C.c_int64 * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
