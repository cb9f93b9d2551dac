"""Synthetic workloads of the BASELINE.json configurations: coupling maps, gatesets built the way
`BaseSynthesisEnv.from_coupling_map` builds them (reference src/qiskit_gym/envs/synthesis.py:91-103),
and seeded random targets in the reference's `set_state` wire format.  Pure numpy, no GPU, no oracle."""
from __future__ import annotations

import numpy as np

from . import _abi

PERM, LF, CLIFF, PAULI = _abi.ENV_PERMUTATION, _abi.ENV_LINEAR_FUNCTION, _abi.ENV_CLIFFORD, _abi.ENV_PAULI_NETWORK
ONE_Q = ("H", "S", "Sdg", "SX", "SXdg")
TWO_Q = ("CX", "CZ", "SWAP")
ALL_GATES = ONE_Q + TWO_Q

# CouplingMap.from_grid(3, 3, bidirectional=False) (examples/intro.ipynb cell 16)
GRID3 = [(0, 1), (0, 3), (1, 2), (1, 4), (2, 5), (3, 4), (3, 6), (4, 5), (4, 7), (5, 8), (6, 7), (7, 8)]
# 27-qubit Falcon heavy-hex coupling map
HEAVY_HEX_27 = [(0, 1), (1, 2), (1, 4), (2, 3), (3, 5), (4, 7), (5, 8), (6, 7), (7, 10), (8, 9), (8, 11), (10, 12), (11, 14),
                (12, 13), (12, 15), (13, 14), (14, 16), (15, 18), (16, 19), (17, 18), (18, 21), (19, 20), (19, 22), (21, 23),
                (22, 25), (23, 24), (24, 25), (25, 26)]


def line_edges(n, bidirectional=True):
    e = []
    for i in range(n - 1):
        e.append((i, i + 1))
        if bidirectional:
            e.append((i + 1, i))
    return sorted(e)


def full_edges(n):
    return sorted((i, j) for i in range(n) for j in range(n) if i != j)


def gateset_from_coupling_map(coupling_map, basis_gates):
    """synthesis.py:91-103: edges sorted; num_qubits = max index + 1; per basis gate, 1-qubit gates over all
    qubits, 2-qubit gates over the sorted edge list.  Returns (num_qubits, gateset)."""
    edges = sorted(tuple(e) for e in coupling_map)
    num_qubits = max(max(e) for e in edges) + 1
    gs = []
    for g in basis_gates:
        if g in ONE_Q:
            gs += [(g, (q,)) for q in range(num_qubits)]
        else:
            assert g in TWO_Q, f"Gate {g} not supported!"
            gs += [(g, e) for e in edges]
    return num_qubits, gs


def baseline_configs():
    """name -> (env kind, num_qubits, gateset, extra constructor kwargs) for BASELINE.json configs[0..4]."""
    out = {}
    n, gs = gateset_from_coupling_map(GRID3, ("SWAP",))
    out["C1_perm_grid3"] = (PERM, n, gs, {})
    n, gs = gateset_from_coupling_map(line_edges(8), ("CX",))
    out["C2_lf8_line"] = (LF, n, gs, {})
    n, gs = gateset_from_coupling_map(full_edges(8), ("H", "S", "CX"))
    out["C3_clifford8_full"] = (CLIFF, n, gs, {})
    n, gs = gateset_from_coupling_map(line_edges(10), ALL_GATES)
    out["C4_pauli10_line"] = (PAULI, n, gs, {"max_rotations": 5})
    n, gs = gateset_from_coupling_map(HEAVY_HEX_27, ("SWAP",))
    out["C5_perm27_heavyhex"] = (PERM, n, gs, {})
    return out


def _scrambled_matrices(rng, kind, n, gateset, B, scramble):
    """identity scrambled by `scramble` uniform gateset actions (reset(), clifford.rs:306-316), vectorised over B."""
    D = 2 * n if kind != LF else n
    M = np.zeros((B, D, D), dtype=np.uint8)
    M[:, np.arange(D), np.arange(D)] = 1
    kinds = np.array([_abi.GATE_NAMES.index(next(g for g in _abi.GATE_NAMES if g.lower() == name.lower())) for name, _ in gateset])
    q0s = np.array([idx[0] for _, idx in gateset])
    q1s = np.array([idx[1] if len(idx) > 1 else 0 for _, idx in gateset])
    rows = np.arange(B)

    def xor(sel, dst, src):
        if sel.any():
            M[rows[sel], dst[sel]] ^= M[rows[sel], src[sel]]

    def swap(sel, a, b):
        if sel.any():
            ra, rb = M[rows[sel], a[sel]].copy(), M[rows[sel], b[sel]].copy()
            M[rows[sel], a[sel]] = rb
            M[rows[sel], b[sel]] = ra

    for _ in range(scramble):
        act = rng.integers(0, len(gateset), size=B)
        k, a, b = kinds[act], q0s[act], q1s[act]
        ne = a != b
        if kind == LF:
            xor((k == 5) & ne, b, a)
            swap((k == 7) & ne, a, b)
        else:
            swap(k == 0, a, n + a)
            xor((k == 1) | (k == 2), n + a, a)
            xor((k == 3) | (k == 4), a, n + a)
            cx = (k == 5) & ne
            xor(cx, b, a); xor(cx, n + a, n + b)
            cz = (k == 6) & ne
            xor(cz, n + a, b); xor(cz, n + b, a)
            sw = (k == 7) & ne
            swap(sw, a, b); swap(sw, n + a, n + b)
    return M


def random_pauli_labels(rng, n, count, min_weight=2):
    out = []
    while len(out) < count:
        lab = "".join(rng.choice(list("IXYZ"), size=n))
        if sum(ch != "I" for ch in lab) >= min(min_weight, n):
            out.append(lab)
    return out


def random_targets(kind, n, gateset, B, seed, scramble=256, num_rotations=5, vary_rotations=False):
    """Seeded `set_state` payloads: Permutation: uniform random permutations; LinearFunction /
    Clifford: identity scrambled by `scramble` random gateset actions; PauliNetwork: a scrambled tableau (H/S/CX
    gates of the gateset) plus `num_rotations` random Pauli labels of weight >= 2.  Returns int64[B, payload_len]
    (PauliNetwork payloads are zero-padded to the longest)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == PERM:
        return np.stack([rng.permutation(n) for _ in range(B)]).astype(np.int64)
    if kind in (LF, CLIFF):
        return _scrambled_matrices(rng, kind, n, gateset, B, scramble).reshape(B, -1).astype(np.int64)
    cl = [g for g in gateset if g[0] in ("H", "S", "CX")] or [("H", (0,))]
    tabs = _scrambled_matrices(rng, CLIFF, n, cl, B, scramble).reshape(B, -1).astype(np.int64)
    payloads = []
    for b in range(B):
        R = int(rng.integers(0, num_rotations + 1)) if vary_rotations else num_rotations
        st = [R] + tabs[b].tolist()
        for lab in random_pauli_labels(rng, n, R):
            st += [len(lab)] + [ord(ch) for ch in lab]
        payloads.append(st)
    stride = max(len(p) for p in payloads)
    arr = np.zeros((B, stride), dtype=np.int64)
    for b, p in enumerate(payloads):
        arr[b, : len(p)] = p
    return arr


def payload_lengths(kind, n, targets):
    """Length of each payload row of random_targets() (PauliNetwork rows are self-delimiting)."""
    B = targets.shape[0]
    if kind != PAULI:
        return np.full(B, targets.shape[1], dtype=np.int64)
    lens = np.zeros(B, dtype=np.int64)
    for b in range(B):
        R = int(targets[b, 0]); i = 1 + 4 * n * n
        for _ in range(R):
            i += 1 + int(targets[b, i])
        lens[b] = i
    return lens


def random_actions(rng, T, B, A, invalid_rate=0.0):
    """int32[T, B] uniform in [0, A); a fraction may be out of range (state no-ops that still tick depth)."""
    a = rng.integers(0, A, size=(T, B), dtype=np.int64)
    if invalid_rate > 0:
        bad = rng.random((T, B)) < invalid_rate
        a = np.where(bad, A + rng.integers(0, 5, size=(T, B)), a)
    return a.astype(np.int32)
