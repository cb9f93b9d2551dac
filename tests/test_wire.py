"""Wire formats either side of the hot path (qiskit_gym_b200/wire.py, specs.py), CPU only.

The encoders are checked against the oracle's env semantics (a target encoded with get_state and then driven with the
target's own gates must end in the solved state — that is what makes the synthesised circuit equal the target, reference
envs/synthesis.py:206-209, 220-224, 254-263); the tableau simulator against a brute-force unitary simulator."""
import itertools
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from qiskit_gym_b200 import wire
from tests import helpers as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# ---- brute-force reference: dense unitaries ------------------------------------------------------------
I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
HG = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
SG = np.array([[1, 0], [0, 1j]], dtype=complex)
SXG = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex)
ONE = {"h": HG, "s": SG, "sdg": SG.conj().T, "sx": SXG, "sxdg": SXG.conj().T, "x": X, "y": Y, "z": Z}


def op_on(n, mats):
    """kron with qubit 0 as the least significant (Qiskit's little-endian convention)"""
    out = np.array([[1]], dtype=complex)
    for q in range(n - 1, -1, -1):
        out = np.kron(out, mats.get(q, I2))
    return out


def two_qubit(n, name, a, b):
    dim = 2 ** n
    U = np.zeros((dim, dim), dtype=complex)
    for s in range(dim):
        ba, bb = (s >> a) & 1, (s >> b) & 1
        if name == "cx":
            t = s ^ (ba << b)
            U[t, s] = 1
        elif name == "cz":
            U[s, s] = -1 if (ba and bb) else 1
        else:  # swap
            t = s & ~((1 << a) | (1 << b)) | (bb << a) | (ba << b)
            U[t, s] = 1
    return U


def unitary(gates, n):
    U = np.eye(2 ** n, dtype=complex)
    for name, qs in gates:
        g = name.lower()
        G = op_on(n, {qs[0]: ONE[g]}) if g in ONE else two_qubit(n, g, qs[0], qs[1])
        U = G @ U
    return U


def pauli_matrix(n, xs, zs, sign):
    mats = {}
    for q in range(n):
        if xs[q] and zs[q]:
            mats[q] = Y
        elif xs[q]:
            mats[q] = X
        elif zs[q]:
            mats[q] = Z
    return (-1 if sign else 1) * op_on(n, mats)


def random_gates(rng, n, k, names=("h", "s", "sdg", "sx", "sxdg", "cx", "cz", "swap")):
    out = []
    for _ in range(k):
        g = names[int(rng.integers(len(names)))]
        if g in ("cx", "cz", "swap"):
            a, b = rng.choice(n, size=2, replace=False)
            out.append((g, (int(a), int(b))))
        else:
            out.append((g, (int(rng.integers(n)),)))
    return out


@pytest.mark.parametrize("n", [1, 2, 3])
def test_tableau_matches_dense_unitaries(n):
    rng = np.random.default_rng(10 + n)
    names = ("h", "s", "sdg", "sx", "sxdg", "x", "y", "z") + (("cx", "cz", "swap") if n > 1 else ())
    for trial in range(20):
        gates = random_gates(rng, n, 12, names)
        T = wire.StabilizerTableau.from_gates(gates, n)
        U = unitary(gates, n)
        for i in range(2 * n):
            xs = np.zeros(n, int); zs = np.zeros(n, int)
            (xs if i < n else zs)[i % n] = 1
            want = U @ pauli_matrix(n, xs, zs, 0) @ U.conj().T
            got = pauli_matrix(n, T.x[i], T.z[i], T.p[i])
            assert np.allclose(want, got, atol=1e-9), (gates, i)


def test_gf2_and_symplectic_inverse():
    rng = np.random.default_rng(3)
    mats = []
    while len(mats) < 40:
        M = rng.integers(0, 2, size=(7, 7), dtype=np.uint8)
        if round(abs(np.linalg.det(M.astype(float)))) % 2 == 1:
            mats.append(M)
    mats = np.stack(mats)
    inv = wire.gf2_inverse(mats)
    assert np.array_equal((mats.astype(int) @ inv.astype(int)) % 2, np.broadcast_to(np.eye(7, dtype=int), (40, 7, 7)))
    assert np.array_equal(wire.gf2_inverse(mats[0]), inv[0])
    with pytest.raises(ValueError):
        wire.gf2_inverse(np.zeros((3, 3), dtype=np.uint8))
    for n in (2, 4):
        F = np.stack([wire.StabilizerTableau.from_gates(random_gates(rng, n, 30), n).symplectic() for _ in range(10)])
        assert np.array_equal(wire.symplectic_inverse(F), wire.gf2_inverse(F))


def test_invert_gates_is_the_inverse_circuit():
    rng = np.random.default_rng(5)
    for _ in range(10):
        g = random_gates(rng, 3, 15)
        T = wire.StabilizerTableau.from_gates(g + wire.invert_gates(g), 3)
        assert np.array_equal(T.to_array(), wire.StabilizerTableau(3).to_array())


def _gateset_all(n):
    return H.gateset_from_coupling_map(H.W.full_edges(n), H.ALL_GATES)[1]


def test_clifford_state_is_solved_by_the_targets_own_gates():
    """get_state(C) then C's gates in order == identity (this is why the solution equals C up to signs)."""
    rng = np.random.default_rng(7)
    n = 4
    gs = _gateset_all(n)
    index = {(g.lower(), tuple(q)): i for i, (g, q) in enumerate(gs)}
    for trial in range(10):
        gates = random_gates(rng, n, 25)
        tab = wire.StabilizerTableau.from_gates(gates, n).to_array()
        state = wire.clifford_state(tab)
        assert np.array_equal(state, wire.clifford_state(tab[:, :-1]))
        env = orc.OracleEnv(H.CLIFF, n, gs, add_inverts=False, add_perms=False)
        env.set_state(state.tolist())
        for g, q in gates:
            env.step(index[(g, tuple(q))])
        assert env.success()
    # batched form
    tabs = np.stack([wire.StabilizerTableau.from_gates(random_gates(rng, n, 9), n).to_array() for _ in range(5)])
    batch = wire.clifford_state(tabs)
    assert batch.shape == (5, 4 * n * n) and np.array_equal(batch[3], wire.clifford_state(tabs[3]))


def test_linear_function_state_is_solved_by_the_targets_own_gates():
    rng = np.random.default_rng(8)
    n = 6
    gs = H.gateset_from_coupling_map(H.W.full_edges(n), ("CX", "SWAP"))[1]
    index = {(g.lower(), tuple(q)): i for i, (g, q) in enumerate(gs)}
    for trial in range(10):
        gates = random_gates(rng, n, 30, ("cx", "swap"))
        M = np.eye(n, dtype=np.uint8)
        for g, (a, b) in gates:
            if g == "cx":
                M[b] ^= M[a]
            else:
                M[[a, b]] = M[[b, a]]
        env = orc.OracleEnv(H.LF, n, gs, add_inverts=False, add_perms=False)
        env.set_state(wire.linear_function_state(M).tolist())
        for g, q in gates:
            env.step(index[(g, tuple(q))])
        assert env.success()
    Ms = np.stack([np.eye(n, dtype=np.uint8)] * 3)
    assert np.array_equal(wire.linear_function_state(Ms), np.stack([np.eye(n, dtype=np.int64).reshape(-1)] * 3))


def test_permutation_state():
    rng = np.random.default_rng(9)
    p = rng.permutation(9)
    s = wire.permutation_state(p)
    assert np.array_equal(s, np.argsort(p)) and np.array_equal(p[s], np.arange(9))
    batch = np.stack([rng.permutation(9) for _ in range(4)])
    assert np.array_equal(wire.permutation_state(batch)[2], np.argsort(batch[2]))
    with pytest.raises(ValueError):
        wire.permutation_state([0, 0, 1])


def test_clifford_phase_fixup_restores_signs():
    rng = np.random.default_rng(11)
    n = 3
    for trial in range(25):
        gates = random_gates(rng, n, 14)
        # target: the same circuit with Pauli gates sprinkled in -> same symplectic matrix, different signs
        target = []
        for g in gates:
            if rng.random() < 0.4:
                target.append((("x", "y", "z")[int(rng.integers(3))], (int(rng.integers(n)),)))
            target.append(g)
        tab = wire.StabilizerTableau.from_gates(target, n).to_array()
        fixed = wire.clifford_phase_fixup(gates, n, tab)
        assert fixed[: len(gates)] == gates and all(g in ("x", "y", "z") for g, _ in fixed[len(gates):])
        assert np.array_equal(wire.StabilizerTableau.from_gates(fixed, n).to_array(), tab)
        U, V = unitary(fixed, n), unitary(target, n)
        k = np.argmax(np.abs(V))
        ph = U.flat[k] / V.flat[k]
        assert abs(abs(ph) - 1) < 1e-9 and np.allclose(U, ph * V, atol=1e-9)
    with pytest.raises(ValueError):
        wire.clifford_phase_fixup([("h", (0,))], 2, wire.StabilizerTableau(2).to_array())


def test_pauli_solution_decoding():
    gs = [("H", (0,)), ("CX", (0, 1)), ("S", (1,))]
    enc = [1, wire.ROTATION_MARKER | (2 << 21) | (1 << 11) | (3 << 1) | 1, 0, wire.ROTATION_MARKER | (0 << 21) | (0 << 11) | (0 << 1)]
    dec = wire.decode_pauli_solution(enc)
    assert dec == [("gate", 1, 0, 0), ("rz", 1, 3, 1), ("gate", 0, 0, 0), ("rx", 0, 0, -1)]
    assert wire.pauli_solution_to_gates(gs, enc) == [("CX", (1, 0)), ("rz", (1,), (3, 1)), ("H", (0,)), ("rx", (0,), (0, -1))]
    g = wire.pauli_solution_to_gates(gs, enc, rotation_params=[0.5, 0.1, 0.2, 0.3])
    assert g[1] == ("rz", (1,), 0.3) and g[3] == ("rx", (0,), -0.5)
    with pytest.raises(Exception):
        wire.pauli_solution_to_gates(gs, enc, rotation_params=[0.5])


def test_pauli_network_state_layout():
    n = 2
    tab = wire.StabilizerTableau.from_gates([("h", (0,)), ("cx", (0, 1))], n).to_array()
    st = wire.pauli_network_state(tab, ["XZ", "IY"])
    assert st[0] == 2 and st[1:17] == tab[:, :-1].T.reshape(-1).astype(int).tolist()
    assert st[17:] == [2, ord("X"), ord("Z"), 2, ord("I"), ord("Y")]
    st2 = wire.pauli_network_state(tab, [], adjoint=True)
    assert st2[1:] == wire.symplectic_inverse(tab[:, :-1]).T.reshape(-1).astype(int).tolist()


def test_from_coupling_map_orders_match_the_notebook():
    """gateset orderings printed by the reference notebook (tests/golden/notebook_kats.json)."""
    from qiskit_gym_b200.specs import SynthSpec
    k = json.load(open(os.path.join(GOLD, "notebook_kats.json")))
    want = [(g, tuple(q)) for g, q in k["perm_grid3_gateset"]]
    p = SynthSpec.from_coupling_map("PermutationEnv", H.W.GRID3, basis_gates=("SWAP",), difficulty=3, metrics_weights={"n_cnots": 1.0})
    assert p.config["gateset"] == want and p.config["num_qubits"] == 9 and p.config["difficulty"] == 3
    assert p.obs_shape() == [9, 9] and p.num_actions() == 12 and p.cls_name == "PermutationEnv"
    want = [(g, tuple(q)) for g, q in k["lf3_gateset"]]
    edges = sorted({tuple(q) for _, q in want})
    assert SynthSpec.from_coupling_map("LinearFunctionEnv", edges, basis_gates=tuple(dict.fromkeys(g for g, _ in want))).config["gateset"] == want
    with pytest.raises(ValueError):
        SynthSpec.from_coupling_map("LinearFunctionEnv", edges, basis_gates=("H",))
    with pytest.raises(ValueError):
        SynthSpec("BogusEnv", {"num_qubits": 2, "gateset": [("CX", (0, 1))]})
    again = SynthSpec.from_json("qiskit_gym.envs.synthesis.LinearFunctionEnv", {"num_qubits": 2, "gateset": [["CX", [0, 1]]], "bogus": 1})
    assert again.config["num_qubits"] == 2 and again.obs_shape() == [2, 2]
