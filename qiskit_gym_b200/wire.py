"""Wire formats either side of the hot path, Qiskit-free and batched.

Encoders: synthesis target -> the `Vec<i64>` payload `Env::set_state` takes
    (reference src/qiskit_gym/envs/synthesis.py: PermutationGym.get_state 254-263, LinearFunctionGym.get_state
    220-224, CliffordGym.get_state 206-209, PauliGym.get_state 414-459), from plain arrays:
      permutation pattern  int[n]            -> argsort(pattern)
      linear function      {0,1}[n, n]       -> inverse over GF(2), row-major
      Clifford tableau     bool[2n, 2n(+1)]  -> symplectic part of the adjoint, transposed, row-major
      Pauli network        tableau + labels  -> [R, tableau..., len, chars...]
    All of them accept a leading batch dimension.
Decoders: `Env::solution` action list -> gate list [(name, qubits)] (rl/synthesis.py:141-147
    `gate_list_to_circuit`, envs/synthesis.py:138-149), the Clifford Pauli-layer phase fix-up
    (envs/synthesis.py:162-177, 211-217) and the PauliNetwork rotation decoding (35-61, 461-500).

`StabilizerTableau` is a small CHP-style simulator with phase bits in Qiskit's Clifford.tableau layout
(rows: destabilisers then stabilisers; columns: x | z | phase), used for the phase fix-up and for building
tableaux from gate lists when Qiskit is not installed.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

ONE_Q_GATES = ["H", "S", "Sdg", "SX", "SXdg"]
TWO_Q_GATES = ["CX", "CZ", "SWAP"]
ROTATION_MARKER = 0x80000000          # pauli.rs ROTATION_MARKER
AXIS_NAMES = ("rx", "ry", "rz")

GateList = List[Tuple[str, Tuple[int, ...]]]


# ------------------------------------------------------------------------------------------------------
# GF(2) linear algebra (batched)
# ------------------------------------------------------------------------------------------------------
def gf2_inverse(mats: np.ndarray) -> np.ndarray:
    """Inverse over GF(2) of one matrix [D, D] or a batch [B, D, D] (Gauss-Jordan, vectorised over the batch).
    Raises ValueError if any matrix is singular."""
    M = np.asarray(mats).astype(np.uint8) & 1
    single = M.ndim == 2
    if single:
        M = M[None]
    B, D, D2 = M.shape
    if D != D2:
        raise ValueError("gf2_inverse needs square matrices")
    A = np.concatenate([M, np.broadcast_to(np.eye(D, dtype=np.uint8), (B, D, D))], axis=2).copy()
    rows = np.arange(B)
    for col in range(D):
        # first row >= col with a one in this column
        cand = A[:, col:, col]
        piv = cand.argmax(axis=1) + col
        if not cand.any(axis=1).all():
            raise ValueError("singular matrix over GF(2)")
        tmp = A[rows, piv].copy()
        A[rows, piv] = A[rows, col]
        A[rows, col] = tmp
        sel = A[:, :, col].copy()
        sel[:, col] = 0
        A ^= sel[:, :, None] * A[:, col][:, None, :]
    out = A[:, :, D:]
    return out[0] if single else out


def symplectic_inverse(F: np.ndarray) -> np.ndarray:
    """Inverse of a symplectic matrix in (x | z) column order: F^-1 = J F^T J with J = [[0, I], [I, 0]]."""
    F = np.asarray(F).astype(np.uint8) & 1
    n = F.shape[-1] // 2
    Ft = np.swapaxes(F, -1, -2)
    # J A J swaps the two row blocks and the two column blocks
    out = np.empty_like(Ft)
    out[..., :n, :n] = Ft[..., n:, n:]
    out[..., :n, n:] = Ft[..., n:, :n]
    out[..., n:, :n] = Ft[..., :n, n:]
    out[..., n:, n:] = Ft[..., :n, :n]
    return out


# ------------------------------------------------------------------------------------------------------
# Stabiliser tableau with phases
# ------------------------------------------------------------------------------------------------------
class StabilizerTableau:
    """Clifford as a tableau with sign bits: row i < n is the image of X_i, row n+i the image of Z_i;
    columns [0, n) are x bits, [n, 2n) z bits, column 2n the sign.  Gates are appended (applied after the
    current operator), which acts on the qubit's columns of every row."""

    def __init__(self, num_qubits: int):
        n = int(num_qubits)
        self.n = n
        self.x = np.zeros((2 * n, n), dtype=np.uint8)
        self.z = np.zeros((2 * n, n), dtype=np.uint8)
        self.p = np.zeros(2 * n, dtype=np.uint8)
        self.x[np.arange(n), np.arange(n)] = 1
        self.z[n + np.arange(n), np.arange(n)] = 1

    # -- constructors -----------------------------------------------------------------------------------
    @classmethod
    def from_array(cls, tableau) -> "StabilizerTableau":
        t = np.asarray(tableau).astype(np.uint8) & 1
        if t.ndim != 2 or t.shape[0] % 2 or t.shape[1] not in (t.shape[0], t.shape[0] + 1):
            raise ValueError("tableau must be [2n, 2n] or [2n, 2n+1]")
        n = t.shape[0] // 2
        out = cls(n)
        out.x = t[:, :n].copy()
        out.z = t[:, n:2 * n].copy()
        out.p = t[:, 2 * n].copy() if t.shape[1] == 2 * n + 1 else np.zeros(2 * n, dtype=np.uint8)
        return out

    @classmethod
    def from_gates(cls, gates: Iterable, num_qubits: int) -> "StabilizerTableau":
        out = cls(num_qubits)
        for name, qubits in gates:
            out.append(name, qubits)
        return out

    def copy(self) -> "StabilizerTableau":
        out = StabilizerTableau(self.n)
        out.x, out.z, out.p = self.x.copy(), self.z.copy(), self.p.copy()
        return out

    def to_array(self, phase: bool = True) -> np.ndarray:
        cols = [self.x, self.z] + ([self.p[:, None]] if phase else [])
        return np.concatenate(cols, axis=1).astype(bool)

    def symplectic(self) -> np.ndarray:
        return np.concatenate([self.x, self.z], axis=1)

    # -- gates -------------------------------------------------------------------------------------------
    def _h(self, q):
        self.p ^= self.x[:, q] & self.z[:, q]
        self.x[:, q], self.z[:, q] = self.z[:, q].copy(), self.x[:, q].copy()

    def _s(self, q):
        self.p ^= self.x[:, q] & self.z[:, q]
        self.z[:, q] ^= self.x[:, q]

    def _sdg(self, q):
        self.p ^= self.x[:, q] & (self.z[:, q] ^ 1)
        self.z[:, q] ^= self.x[:, q]

    def _x(self, q):
        self.p ^= self.z[:, q]

    def _z(self, q):
        self.p ^= self.x[:, q]

    def _y(self, q):
        self.p ^= self.x[:, q] ^ self.z[:, q]

    def _cx(self, c, t):
        self.p ^= self.x[:, c] & self.z[:, t] & (self.x[:, t] ^ self.z[:, c] ^ 1)
        self.x[:, t] ^= self.x[:, c]
        self.z[:, c] ^= self.z[:, t]

    def append(self, name: str, qubits: Sequence[int]) -> "StabilizerTableau":
        g = name.strip().lower()
        q = [int(v) for v in qubits]
        if any(v < 0 or v >= self.n for v in q):
            raise ValueError(f"gate {name} on qubits {q} outside the register")
        if g == "h":
            self._h(q[0])
        elif g == "s":
            self._s(q[0])
        elif g == "sdg":
            self._sdg(q[0])
        elif g == "sx":          # SX = H S H up to a global phase
            self._h(q[0]); self._s(q[0]); self._h(q[0])
        elif g == "sxdg":
            self._h(q[0]); self._sdg(q[0]); self._h(q[0])
        elif g == "x":
            self._x(q[0])
        elif g == "y":
            self._y(q[0])
        elif g == "z":
            self._z(q[0])
        elif g == "cx":
            self._cx(q[0], q[1])
        elif g == "cz":
            self._h(q[1]); self._cx(q[0], q[1]); self._h(q[1])
        elif g == "swap":
            self._cx(q[0], q[1]); self._cx(q[1], q[0]); self._cx(q[0], q[1])
        else:
            raise TypeError(f"Gate {name} on qubits {q} not supported.")
        return self


def invert_gates(gates: Iterable) -> GateList:
    """Inverse circuit as a gate list (reversed order, S <-> Sdg, SX <-> SXdg)."""
    inv = {"s": "sdg", "sdg": "s", "sx": "sxdg", "sxdg": "sx"}
    out = []
    for name, qubits in reversed(list(gates)):
        g = name.strip().lower()
        out.append((inv.get(g, g), tuple(int(q) for q in qubits)))
    return out


# ------------------------------------------------------------------------------------------------------
# Encoders (targets -> set_state payloads)
# ------------------------------------------------------------------------------------------------------
def permutation_state(pattern) -> np.ndarray:
    """PermutationGym.get_state (envs/synthesis.py:254-263): the inverse permutation, `argsort(pattern)`.
    pattern: int[n] or int[B, n]."""
    p = np.asarray(pattern, dtype=np.int64)
    if p.ndim not in (1, 2):
        raise ValueError("permutation pattern must be [n] or [B, n]")
    n = p.shape[-1]
    if not np.array_equal(np.sort(p, axis=-1), np.broadcast_to(np.arange(n), p.shape)):
        raise ValueError("not a permutation of 0..n-1")
    return np.argsort(p, axis=-1, kind="stable").astype(np.int64)


def linear_function_state(matrix) -> np.ndarray:
    """LinearFunctionGym.get_state (envs/synthesis.py:220-224): the linear matrix of the inverse circuit,
    row-major.  matrix: {0,1}[n, n] or [B, n, n]  ->  int64[n*n] or [B, n*n]."""
    inv = gf2_inverse(matrix)
    return inv.reshape(inv.shape[:-2] + (-1,)).astype(np.int64)


def clifford_state(tableau) -> np.ndarray:
    """CliffordGym.get_state (envs/synthesis.py:206-209): `adjoint().tableau[:, :-1].T.flatten()`.
    tableau: bool[2n, 2n] / [2n, 2n+1] (Qiskit Clifford.tableau layout; the phase column is ignored here) or a
    batch [B, 2n, 2n(+1)]  ->  int64[4n^2] or [B, 4n^2]."""
    t = np.asarray(tableau).astype(np.uint8) & 1
    D = t.shape[-2]
    if t.shape[-1] not in (D, D + 1) or D % 2:
        raise ValueError("tableau must be [2n, 2n] or [2n, 2n+1]")
    F = t[..., :D]
    adj_t = np.swapaxes(symplectic_inverse(F), -1, -2)
    return adj_t.reshape(adj_t.shape[:-2] + (-1,)).astype(np.int64)


def pauli_network_state(tableau, rotations: Sequence[str], adjoint: bool = False) -> List[int]:
    """PauliGym.get_state (envs/synthesis.py:414-459) for a (tableau, rotation labels) pair:
    `[R, tableau[:, :-1].T.flatten()..., len, chars..., ...]`.  `adjoint=True` first takes the adjoint of the
    tableau (the reference does that for raw Clifford / circuit inputs, not for tuple inputs)."""
    t = np.asarray(tableau).astype(np.uint8) & 1
    D = t.shape[0]
    F = t[:, :D]
    if adjoint:
        F = symplectic_inverse(F)
    state = [len(rotations)] + F.T.reshape(-1).astype(np.int64).tolist()
    for rot in rotations:
        state.append(len(rot))
        state.extend(ord(c) for c in rot)
    return state


# ------------------------------------------------------------------------------------------------------
# Decoders (solutions -> gate lists)
# ------------------------------------------------------------------------------------------------------
def solution_to_gates(gateset: Sequence, actions: Iterable[int]) -> GateList:
    """`[gateset[a] for a in actions]` (envs/synthesis.py:145-147) with normalised tuples."""
    out = []
    for a in actions:
        name, qubits = gateset[int(a)]
        out.append((str(name), tuple(int(q) for q in qubits)))
    return out


def clifford_phase_fixup(gates: GateList, num_qubits: int, target_tableau) -> GateList:
    """CliffordGym.post_process_synthesis (envs/synthesis.py:162-177, 211-217): the synthesised gates reproduce
    the target's symplectic matrix; one trailing layer of Pauli gates fixes the signs.  Returns gates + layer.
    With C the target and G the synthesised circuit, G^-1 then C is a Pauli operator whose tableau signs say, per qubit:
    destabiliser and stabiliser sign -> Y, stabiliser only -> X, destabiliser only -> Z."""
    tgt = StabilizerTableau.from_array(target_tableau)
    got = StabilizerTableau.from_gates(gates, num_qubits)
    if not np.array_equal(tgt.symplectic(), got.symplectic()):
        raise ValueError("the gate list does not implement the target Clifford (symplectic parts differ)")
    diff = (tgt.p ^ got.p).astype(np.uint8)
    # C = G then P: row i of G anticommutes with P iff diff[i]; solve F v = diff, v = (pz | px)
    v = (gf2_inverse(got.symplectic()).astype(np.uint8) @ diff) & 1
    n = num_qubits
    pz, px = v[:n], v[n:]
    layer = []
    for q in range(n):
        if px[q] and pz[q]:
            layer.append(("y", (q,)))
        elif px[q]:
            layer.append(("x", (q,)))
        elif pz[q]:
            layer.append(("z", (q,)))
    return list(gates) + layer


def decode_pauli_solution(encoded_solution: Iterable[int]):
    """PauliNetwork solution entries (pauli.rs:698-716; Python decoder envs/synthesis.py:35-61):
    plain gate actions, or ROTATION_MARKER | axis << 21 | qubit << 11 | index << 1 | phase bit.
    -> [("gate", action, 0, 0) | ("rx"|"ry"|"rz", qubit, index, +1|-1)]"""
    out = []
    for val in encoded_solution:
        val = int(val)
        if val >= ROTATION_MARKER:
            out.append((AXIS_NAMES[(val >> 21) & 0x3], (val >> 11) & 0x3FF, (val >> 1) & 0x3FF, 1 if (val & 1) else -1))
        else:
            out.append(("gate", val, 0, 0))
    return out


def pauli_solution_to_gates(gateset: Sequence, encoded_solution: Iterable[int], rotation_params: Sequence[float] | None = None):
    """PauliGym._reconstruct_circuit_from_solution (envs/synthesis.py:461-500) without the final Clifford phase
    correction (that one needs a Clifford synthesiser; the caller gets the gate list and may append its own).
    CX qubits are emitted reversed (the PauliNetwork cnot convention, pauli_network.rs:196-207).  Rotations come
    out as ("rx"|"ry"|"rz", (qubit,), angle) when `rotation_params` is given, else ("rx", (qubit,), (index, sign))."""
    out = []
    for kind, a1, a2, a3 in decode_pauli_solution(encoded_solution):
        if kind == "gate":
            name, qubits = gateset[a1]
            qs = tuple(int(q) for q in qubits)
            if name.lower() == "cx":
                qs = qs[::-1]
            out.append((str(name), qs))
        else:
            if rotation_params is not None:
                if a2 >= len(rotation_params):
                    raise Exception("Too few rotation parameters stored for synthesis!")
                out.append((kind, (a1,), a3 * rotation_params[a2]))
            else:
                out.append((kind, (a1,), (a2, a3)))
    return out


# ------------------------------------------------------------------------------------------------------
# Optional Qiskit bridges (only if qiskit is importable; never required by the engine)
# ------------------------------------------------------------------------------------------------------
def have_qiskit() -> bool:
    try:
        import qiskit  # noqa: F401
        return True
    except Exception:
        return False


def gates_to_circuit(gates: Iterable, num_qubits: int):
    """gate list -> qiskit.QuantumCircuit (rl/synthesis.py:141-147); ImportError without Qiskit."""
    from qiskit import QuantumCircuit
    qc = QuantumCircuit(num_qubits)
    for g in gates:
        name, qubits = g[0], g[1]
        if len(g) == 3:
            getattr(qc, name.lower())(g[2], *qubits)
        else:
            getattr(qc, name.lower())(*qubits)
    return qc
