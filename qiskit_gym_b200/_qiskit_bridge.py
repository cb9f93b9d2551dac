"""PauliGym <-> Qiskit objects (only imported when Qiskit is installed; the engine never needs it).

The array / gate-list forms in wire.py cover everything else; what is left here needs Qiskit's own algebra:
parsing a QuantumCircuit with rx/ry/rz into (Clifford, evolved rotation labels, angles)
(reference src/qiskit_gym/envs/synthesis.py:320-361) and the final Clifford phase correction that calls
`Clifford.to_circuit()` (synthesis.py:492-498).  Untested in this image (no Qiskit wheel available).
"""
from __future__ import annotations

import numpy as np


def _parse(circuit):
    """synthesis.py:320-361: walk the circuit once; Clifford gates accumulate into `cliff`, each rotation becomes the
    adjoint label of its single-qubit Pauli evolved through the Clifford seen so far."""
    from qiskit.exceptions import QiskitError
    from qiskit.quantum_info import Clifford, Pauli

    n = circuit.num_qubits
    cliff = Clifford(np.eye(2 * n, dtype=bool))
    labels, angles = [], []
    for inst in circuit.data:
        name = inst.operation.name.lower()
        qubits = [circuit.find_bit(q).index for q in inst.qubits]
        if name in ("rx", "ry", "rz"):
            chars = ["I"] * n
            chars[n - 1 - qubits[0]] = name[1].upper()
            labels.append(Pauli("".join(chars)).evolve(cliff).adjoint().to_label())
            angles.extend(inst.operation.params)
        else:
            try:
                cliff = cliff.compose(inst.operation, qubits)
            except QiskitError:
                raise TypeError(f"Gate {name} on qubits {qubits} not supported.")
    return cliff, labels, angles


def pauli_get_state(env, input, rotations=None):
    """synthesis.py:414-459 for Qiskit inputs."""
    from qiskit import QuantumCircuit
    from qiskit.quantum_info import Clifford

    if isinstance(input, tuple):
        cliff, rotations = input
        env._rotation_params, env._original_circuit = [], None
    elif isinstance(input, QuantumCircuit):
        parsed, rotations, angles = _parse(input)
        cliff = parsed.adjoint()
        env._rotation_params, env._original_circuit = angles, input
    elif isinstance(input, Clifford):
        cliff = input.adjoint()
        rotations = rotations or []
        env._rotation_params, env._original_circuit = [], None
    else:
        raise ValueError(f"Unsupported input type: {type(input)}")
    state = [len(rotations)] + cliff.tableau[:, :-1].T.flatten().astype(int).tolist()
    for rot in rotations:
        state.append(len(rot))
        state.extend(ord(c) for c in rot)
    return state


def pauli_reconstruct(env, full_solution, input):
    """synthesis.py:461-500: gates (CX reversed) + rotations with their stored angles, then the Clifford that makes the
    result equal to the original circuit's Clifford part."""
    from qiskit import QuantumCircuit
    from qiskit.quantum_info import Clifford

    n = env.config["num_qubits"]
    qc = QuantumCircuit(n)
    for kind, a1, a2, a3 in full_solution:
        if kind == "gate":
            name, qubits = env.config["gateset"][a1]
            qubits = list(qubits)[::-1] if name.lower() == "cx" else list(qubits)
            getattr(qc, name.lower())(*qubits)
        else:
            if a2 >= len(env._rotation_params):
                raise Exception("Too few rotation parameters stored for synthesis!")
            getattr(qc, kind)(a3 * env._rotation_params[a2], a1)
    original = input if isinstance(input, QuantumCircuit) else env._original_circuit
    if original is not None:
        rest = qc.inverse().compose(original)
        only_clifford = QuantumCircuit.copy_empty_like(rest)
        for g in rest:
            if g.operation.name not in {"rx", "ry", "rz"}:
                only_clifford.append(g.operation, g.qubits)
        qc = qc.compose(Clifford(only_clifford).to_circuit())
    return qc
