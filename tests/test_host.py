"""CPU tests of the host side of the product library: the C-ABI library loads and exports every symbol the
header declares, gate parsing / validation errors mirror the reference, twists match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from qiskit_gym_b200 import _abi, host
from qiskit_gym_b200 import workloads as W
from qiskit_gym_b200._lib import SYMBOLS, lib
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qg_engine.h")).read()
    declared = re.findall(r"QG_API\s+[\w\s\*]+?\b(qg_\w+)\s*\(", hdr)
    assert len(declared) >= 35
    L = lib()
    for name in declared:
        assert hasattr(L, name), f"libqg_engine.so does not export {name}"
    assert sorted(set(declared)) == sorted(set(SYMBOLS))
    assert b"sm_100a" in L.qg_version()


def test_config_struct_layout_matches_header():
    cfg = _abi.QgConfig()
    lib().qg_config_default(C.byref(cfg), _abi.ENV_PAULI_NETWORK)
    assert (cfg.difficulty, cfg.depth_slope, cfg.max_depth) == (1, 2, 128)
    assert np.float32(cfg.w_n_cnots) == np.float32(0.01) and np.float32(cfg.w_n_gates) == np.float32(0.0001)
    assert (cfg.add_inverts, cfg.add_perms, cfg.track_solution) == (0, 1, 1)
    assert (cfg.max_rotations, cfg.pauli_diff_scale, cfg.final_pauli_layers) == (5, 8, -1)
    assert np.float32(cfg.num_qubits_decay) == np.float32(0.5) and np.float32(cfg.pauli_layer_reward) == np.float32(0.01)


def test_gate_parser_mirrors_common_rs():
    k = lib().qg_gate_kind_from_name
    for i, name in enumerate(_abi.GATE_NAMES):
        assert k(name.encode(), 2 if i >= 5 else 1) == i
        assert k(f"  {name.lower()} ".encode(), 2 if i >= 5 else 1) == i       # trimmed, case-insensitive
        assert k(name.upper().encode(), 1 if i >= 5 else 2) == _abi.QG_ERR_STATE  # wrong arity
    assert k(b"ccx", 3) == _abi.QG_ERR_INVALID
    with pytest.raises(ValueError, match="Unknown gate name"):
        host.make_config(_abi.ENV_CLIFFORD, 2, [("T", (0,))])
    with pytest.raises(ValueError, match="expects 2 indices, got 1"):
        host.make_config(_abi.ENV_CLIFFORD, 2, [("CX", (0,))])
    with pytest.raises(ValueError, match="expects 1 index, got 2"):
        host.make_config(_abi.ENV_CLIFFORD, 2, [("h", (0, 1))])
    with pytest.raises(ValueError, match="exactly 2 items"):
        host.make_config(_abi.ENV_CLIFFORD, 2, [("H", (0,), 1)])
    with pytest.raises(TypeError):
        host.make_config(_abi.ENV_CLIFFORD, 2, [("H", (-1,))])
    with pytest.raises(TypeError):
        host.make_config(_abi.ENV_CLIFFORD, 2, [(3, (0,))])
    with pytest.raises(TypeError):
        host.make_config(_abi.ENV_CLIFFORD, 2, ["H0"])


def test_validation_and_limits():
    cfg = host.make_config(_abi.ENV_CLIFFORD, 2, [("CX", (0, 2))])
    with pytest.raises(ValueError, match="out of range"):
        host.validate(cfg)
    with pytest.raises(NotImplementedError):
        host.validate(host.make_config(_abi.ENV_CLIFFORD, 33, [("H", (0,))]))
    with pytest.raises(NotImplementedError):
        host.validate(host.make_config(_abi.ENV_LINEAR_FUNCTION, 65, [("CX", (0, 1))]))
    with pytest.raises(NotImplementedError):
        host.validate(host.make_config(_abi.ENV_PAULI_NETWORK, 30, [("CX", (0, 1))], max_rotations=5))
    host.validate(host.make_config(_abi.ENV_PERMUTATION, 200, [("SWAP", (0, 199))]))
    assert host.obs_shape(host.make_config(_abi.ENV_PAULI_NETWORK, 10, [("CX", (0, 1))], max_rotations=5)) == [20, 25]
    assert host.obs_shape(host.make_config(_abi.ENV_CLIFFORD, 8, [("H", (0,))])) == [16, 16]
    assert host.obs_shape(host.make_config(_abi.ENV_PERMUTATION, 27, [("SWAP", (0, 1))])) == [27, 27]


def test_metrics_weights_from_hashmap_semantics():
    cfg = host.make_config(_abi.ENV_CLIFFORD, 2, [("H", (0,))], metrics_weights={"n_cnots": 0.5, "bogus": 3.0})
    assert np.float32(cfg.w_n_cnots) == np.float32(0.5) and np.float32(cfg.w_n_gates) == np.float32(0.0001)


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C5_perm27_heavyhex", "clifford3_allgates", "clifford5_allgates", "lf5_line_swap", "perm5_mixed"])
def test_twists_match_oracle(name):
    kind, n, gs, kw = H.config_table()[name]
    cfg = host.make_config(kind, n, gs, **kw)
    obs, act = host.twists(cfg)
    o2, a2 = orc.OracleEnv(kind, n, gs, **kw).twists()
    assert obs == o2 and act == a2 and len(obs) >= 1


def test_twists_special_cases():
    # no two-qubit gate: all permutations in Heap's order (symmetry.rs:84-113), filtered by the gateset
    gs = [("H", (q,)) for q in range(4)] + [("S", (q,)) for q in range(4)]
    obs, act = host.twists(host.make_config(_abi.ENV_CLIFFORD, 4, gs))
    o2, a2 = orc.OracleEnv(H.CLIFF, 4, gs).twists()
    assert len(obs) == 24 and obs == o2 and act == a2
    # all-to-all: every permutation is an automorphism (sorted order)
    n, gs = W.gateset_from_coupling_map(W.full_edges(5), ("CX",))
    obs, act = host.twists(host.make_config(_abi.ENV_LINEAR_FUNCTION, n, gs))
    o2, a2 = orc.OracleEnv(H.LF, n, gs).twists()
    assert len(obs) == 120 and obs == o2 and act == a2
    # duplicate SWAP quirk: later duplicate wins (symmetry.rs:217-223)
    gs = [("SWAP", (0, 1)), ("SWAP", (1, 0)), ("SWAP", (1, 2)), ("SWAP", (2, 1))]
    obs, act = host.twists(host.make_config(_abi.ENV_LINEAR_FUNCTION, 3, gs))
    assert act[0] == [1, 1, 3, 3]
    # add_perms=False and PauliNetwork: empty twists
    assert host.twists(host.make_config(_abi.ENV_LINEAR_FUNCTION, 3, gs, add_perms=False)) == ([], [])
    n, pg = W.gateset_from_coupling_map(W.line_edges(4), W.ALL_GATES)
    assert host.twists(host.make_config(_abi.ENV_PAULI_NETWORK, n, pg, max_rotations=3)) == ([], [])


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qiskit_gym_b200 import BatchedEnv
    kind, n, gs, kw = H.config_table()["C2_lf8_line"]
    with pytest.raises(RuntimeError):
        BatchedEnv(kind, n, gs, 4)
    h = C.c_void_p()
    cfg = host.make_config(kind, n, gs)
    assert lib().qg_create(C.byref(cfg), 0, 4, None, C.byref(h)) == _abi.QG_ERR_CUDA


def test_workload_generators_are_seeded_and_valid():
    for name, (kind, n, gs, kw) in W.baseline_configs().items():
        a = W.random_targets(kind, n, gs, 5, 42, scramble=16)
        b = W.random_targets(kind, n, gs, 5, 42, scramble=16)
        assert np.array_equal(a, b)
        lens = W.payload_lengths(kind, n, a)
        env = orc.OracleEnv(kind, n, gs, add_perms=False)
        for i in range(5):
            env.set_state(a[i, : lens[i]].tolist())   # loads without error
            if kind in (H.LF, H.CLIFF):
                # scrambled matrices stay invertible: reachable from identity by gates
                m = np.array(env.raw_state()).reshape(int(np.sqrt(len(env.raw_state()))), -1)
                assert round(abs(np.linalg.det(m.astype(float)))) % 2 == 1


def _gf2_probe():
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_gf2_probe.so")
    srcs = [os.path.join(here, "gf2_probe.cpp"), os.path.join(here, "..", "qiskit_gym_b200", "csrc", "qg_gf2.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, srcs[0]])
    return C.CDLL(out)


def _pack_rows(M):
    D = M.shape[0]
    bits = M.reshape(-1).astype(np.uint8)
    words = np.zeros((D * D + 31) // 32 + 1, dtype=np.uint32)
    for i in np.nonzero(bits)[0]:
        words[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    return words


def _unpack_rows(words, D):
    return np.array([(int(words[i >> 5]) >> (i & 31)) & 1 for i in range(D * D)], dtype=np.uint8).reshape(D, D)


def test_register_gauss_jordan_and_symplectic_inverse_match_numpy():
    """csrc/qg_gf2.cuh (the add_inverts inverse of LinearFunction / Clifford, linear_function.rs:124-146, clifford.rs:147-170)
    compiled for the host: equal to an independent numpy Gauss-Jordan for every size a bucket covers; singular matrices are
    reported and left untouched."""
    from qiskit_gym_b200 import wire
    P = _gf2_probe()
    rng = np.random.default_rng(2)
    for dmax, sizes in ((8, (1, 2, 5, 8)), (16, (3, 9, 12, 16)), (32, (7, 17, 24, 31, 32))):
        for D in sizes:
            done = 0
            while done < 12:
                M = rng.integers(0, 2, size=(D, D), dtype=np.uint8)
                w = _pack_rows(M); w0 = w.copy()
                rc = P.probe_gauss_jordan(dmax, D, w.ctypes.data_as(C.c_void_p))
                try:
                    want = wire.gf2_inverse(M)
                except ValueError:
                    assert rc == 0 and np.array_equal(w, w0)
                    if D > 4:
                        continue
                    done += 1
                    continue
                assert rc == 1 and np.array_equal(_unpack_rows(w, D), want), (dmax, D)
                done += 1
    for dmax, ns in ((8, (1, 2, 4)), (16, (3, 5, 8)), (32, (6, 11, 16))):
        for n in ns:
            for _ in range(8):
                gates = []
                for _k in range(6 * n):
                    g = ("h", "s", "cx")[int(rng.integers(3))]
                    if g == "cx" and n > 1:
                        a, b = rng.choice(n, size=2, replace=False)
                        gates.append((g, (int(a), int(b))))
                    elif g != "cx":
                        gates.append((g, (int(rng.integers(n)),)))
                F = wire.StabilizerTableau.from_gates(gates, n).symplectic()
                w = _pack_rows(F)
                assert P.probe_symplectic(dmax, n, w.ctypes.data_as(C.c_void_p)) == 1
                assert np.array_equal(_unpack_rows(w, 2 * n), wire.gf2_inverse(F)), (dmax, n)
