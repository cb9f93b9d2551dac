"""Shared builders for the tests: BASELINE.json configurations plus small/odd-sized cases."""
from __future__ import annotations

from oracle import oracle as orc
from qiskit_gym_b200 import workloads as W
from qiskit_gym_b200.workloads import (ALL_GATES, CLIFF, LF, PAULI, PERM, gateset_from_coupling_map, line_edges,  # noqa: F401
                                       payload_lengths, random_actions, random_targets)


def config_table():
    t = dict(W.baseline_configs())

    def add(name, kind, edges, basis, **kw):
        n, gs = gateset_from_coupling_map(edges, basis)
        t[name] = (kind, n, gs, kw)

    add("lf5_line_swap", LF, line_edges(5), ("CX", "SWAP"))
    add("lf11_line", LF, line_edges(11), ("CX", "SWAP"))
    add("lf40_line", LF, line_edges(40), ("CX", "SWAP"))
    add("clifford3_allgates", CLIFF, line_edges(3), ALL_GATES)
    add("clifford5_allgates", CLIFF, line_edges(5), ALL_GATES)
    add("clifford20_line", CLIFF, line_edges(20), ("H", "S", "SX", "CX", "CZ", "SWAP"))
    # power-of-two row widths (the one-word row fast path of the LinearFunction / Clifford gates), every gate kind
    add("lf4_swap", LF, line_edges(4), ("CX", "SWAP"))
    add("lf8_swap", LF, line_edges(8), ("CX", "SWAP"))
    add("lf32_line", LF, line_edges(32), ("CX", "SWAP"))
    add("clifford4_allgates", CLIFF, line_edges(4), ALL_GATES)
    add("clifford8_allgates", CLIFF, line_edges(8), ALL_GATES)
    add("clifford16_allgates", CLIFF, line_edges(16), ALL_GATES)
    add("pauli3_line", PAULI, line_edges(3), ALL_GATES, max_rotations=3)
    add("pauli6_line", PAULI, line_edges(6), ALL_GATES, max_rotations=4, final_pauli_layers=8)
    t["perm5_mixed"] = (PERM, 5, [("SWAP", (0, 1)), ("H", (2,)), ("CX", (1, 2)), ("SWAP", (3, 4)), ("SWAP", (2, 3)), ("CZ", (0, 4))], {})
    return t


def make_cfg(kind, n, gateset, **kw):
    return orc.make_config(kind, n, gateset, **kw)
