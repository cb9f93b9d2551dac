"""CPU tests of the oracle (test infrastructure): golden vectors from the reference's notebook, hand-derived
known answers, the independent Python restatement, and invariants."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref
from tests import helpers as H

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NB = json.load(open(os.path.join(G, "notebook_kats.json")))
DK = json.load(open(os.path.join(G, "derived_kats.json")))


def bits(x):
    return pyref.f32_bits(x)


def succ(env):
    return env.success() if callable(env.success) else env.success


def dense(env):
    shp = env.obs_shape()
    d = np.zeros(shp[0] * shp[1], dtype=np.int64)
    d[env.observe()] = 1
    return d.reshape(shp).tolist()


# ------------------------------------------------------------------ notebook (reference-produced) vectors
def lf3_env(cls=None):
    gs = [(g, tuple(q)) for g, q in NB["lf3_gateset"]]
    if cls is None:
        return orc.OracleEnv(H.LF, 3, gs, add_inverts=False, add_perms=False)
    return cls(3, gs, add_inverts=False)


def test_notebook_gateset_orders():
    n, gs = H.gateset_from_coupling_map(H.line_edges(3), ("CX", "SWAP"))
    assert n == 3 and [[g, list(q)] for g, q in gs] == NB["lf3_gateset"]
    n, gs = H.gateset_from_coupling_map(H.W.GRID3, ("SWAP",))
    assert n == 9 and [[g, list(q)] for g, q in gs] == NB["perm_grid3_gateset"]


@pytest.mark.parametrize("impl", ["cpp", "py"])
def test_notebook_lf_sequences(impl):
    for key in ("lf3_sequence_a", "lf3_sequence_b"):
        seq = NB[key]
        env = lf3_env() if impl == "cpp" else lf3_env(pyref.LinearFunctionRef)
        env.set_state(np.array(seq["start"]).reshape(-1).tolist())
        assert dense(env) == NB["lf3_state_after_set_state"]
        assert env.num_actions() == NB["lf3_action_space"] if impl == "cpp" else True
        assert env.obs_shape() == NB["lf3_obs_space"]
        for a, st, fin in zip(seq["actions"], seq["states"], seq["is_final"]):
            env.step(a)
            assert dense(env) == st
            assert env.is_final() == fin
    env = lf3_env()
    env.set_state(np.array(NB["lf3_state_after_set_state"]).reshape(-1).tolist())
    env.step(NB["lf3_step2"]["action"])
    assert dense(env) == NB["lf3_step2"]["obs"] and env.is_final() == NB["lf3_step2"]["is_final"]


def test_notebook_reset_is_one_gate_from_identity():
    # cell 4: difficulty 1 reset shows identity with exactly one gateset action applied
    gs = [(g, tuple(q)) for g, q in NB["lf3_gateset"]]
    reachable = []
    for a in range(len(gs)):
        env = lf3_env()
        env.set_state(np.eye(3, dtype=int).reshape(-1).tolist())
        env.step(a)
        reachable.append(dense(env))
    assert NB["lf3_reset_difficulty1_obs"] in reachable
    env = orc.OracleEnv(H.LF, 3, gs, difficulty=1, add_inverts=False, add_perms=False)
    for seed in range(20):
        env.reset(seed=seed)
        assert dense(env) in reachable and env.depth() == 2


# ------------------------------------------------------------------ hand-derived vectors
@pytest.mark.parametrize("impl", ["cpp", "py"])
def test_derived_penalties_and_metrics(impl):
    tr = DK["metrics_trace_n3"]
    gs = [(g, tuple(q)) for g, q in tr["gates"]]
    if impl == "cpp":
        env = orc.OracleEnv(H.CLIFF, 3, gs, add_inverts=False, add_perms=False)
        env.set_state(np.eye(6, dtype=int).reshape(-1).tolist())
        for a, want in enumerate(tr["counts_cnots_cxlayers_layers_gates"]):
            env.step(a)
            assert env.counts() == want
    pb = DK["penalty_bits"]
    mk = (lambda gs: orc.OracleEnv(H.CLIFF, 2, gs, add_inverts=False, add_perms=False)) if impl == "cpp" else (lambda gs: pyref.CliffordRef(2, gs, add_inverts=False))
    for name, gate in (("CX", ("CX", (0, 1))), ("SWAP", ("SWAP", (0, 1))), ("CZ", ("CZ", (0, 1))), ("1Q", ("H", (0,)))):
        env = mk([gate])
        st = np.eye(4, dtype=int); st[0, 1] = 1   # never solved by these gates
        env.set_state(st.reshape(-1).tolist())
        env.step(0)
        assert not succ(env)
        assert bits(-env.reward()) == int(pb[name]["penalty"], 16), name
    # solved rewards: apply the gate to its own inverse image
    for name, gate, pre in (("CX", ("CX", (0, 1)), None), ("SWAP", ("SWAP", (0, 1)), None), ("1Q", ("H", (0,)), None)):
        env = mk([gate])
        env.set_state(np.eye(4, dtype=int).reshape(-1).tolist())
        env.step(0)                      # now one gate away from identity (all three are involutions on the tableau)
        state = np.array(env.raw_state()).reshape(-1).tolist()
        env.set_state(state)
        env.step(0)
        assert succ(env)
        assert bits(env.reward()) == int(pb[name]["reward_solved"], 16), name


@pytest.mark.parametrize("impl", ["cpp", "py"])
def test_derived_perm_clifford_pauli(impl):
    k = DK["perm_n3"]
    gs = [(g, tuple(q)) for g, q in k["gateset"]]
    env = orc.OracleEnv(H.PERM, 3, gs, add_inverts=False, add_perms=False) if impl == "cpp" else pyref.PermutationRef(3, gs, add_inverts=False)
    env.set_state(k["set_state"])
    assert env.observe() == k["observe0"]
    for s in k["steps"]:
        env.step(s["action"])
        assert list(env.raw_state()) == s["state"] and env.observe() == s["observe"]
        assert bits(env.reward()) == int(s["reward"], 16) and env.is_final() == s["final"]
    assert env.masks() == k["masks_end"] and env.solution() == k["solution"]

    k = DK["clifford_n2"]
    gs = [(g, tuple(q)) for g, q in k["gateset"]]
    env = orc.OracleEnv(H.CLIFF, 2, gs, add_inverts=False, add_perms=False) if impl == "cpp" else pyref.CliffordRef(2, gs, add_inverts=False)
    env.set_state(np.eye(4, dtype=int).reshape(-1).tolist())
    for a, want in zip(k["actions"], k["observe"]):
        env.step(a)
        assert env.observe() == want

    k = DK["pauli_n3"]
    gs = [(g, tuple(q)) for g, q in k["gateset"]]
    env = orc.OracleEnv(H.PAULI, 3, gs, max_rotations=k["max_rotations"], add_perms=False) if impl == "cpp" else pyref.PauliRef(3, gs, max_rotations=k["max_rotations"])
    st = [len(k["rotations"])] + np.eye(6, dtype=int).reshape(-1).tolist()
    for lab in k["rotations"]:
        st += [len(lab)] + [ord(c) for c in lab]
    env.set_state(st)
    assert env.obs_shape() == k["obs_shape"] and env.observe() == k["observe0"]
    for s in k["steps"]:
        env.step(s["action"])
        assert env.solution() == s["solution"] and env.observe() == s["observe"]
        assert bits(env.reward()) == int(s["reward"], 16)


def test_derived_symmetry():
    k = DK["symmetry"]
    tab = H.config_table()
    kind, n, gs, _ = tab["C1_perm_grid3"]
    obs, act = orc.OracleEnv(kind, n, gs).twists()
    assert len(obs) == k["grid3_automorphisms"] and all(sorted(p) == list(range(81)) for p in obs)
    kind, n, gs, _ = tab["C5_perm27_heavyhex"]
    obs, act = orc.OracleEnv(kind, n, gs).twists()
    assert len(obs) == k["heavy_hex27_automorphisms"]
    gs = [(g, tuple(q)) for g, q in k["duplicate_swap_gateset"]]
    obs, act = orc.OracleEnv(H.LF, 3, gs).twists()
    assert act[0] == k["duplicate_swap_identity_act_perm"]


# ------------------------------------------------------------------ C++ oracle vs independent Python restatement
@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "clifford3_allgates", "lf5_line_swap",
                                  "perm5_mixed", "pauli3_line", "pauli6_line", "C4_pauli10_line"])
@pytest.mark.parametrize("inverts", [False, True])
def test_cpp_oracle_matches_python_restatement(name, inverts):
    kind, n, gs, kw = H.config_table()[name]
    if kind == H.PAULI and inverts:
        pytest.skip("PauliNetwork has no add_inverts")
    rng = np.random.Generator(np.random.PCG64(11))
    B, T = 6, 30
    targets = H.random_targets(kind, n, gs, B, 5, scramble=12, num_rotations=kw.get("max_rotations", 5) + 1, vary_rotations=True)
    lens = H.payload_lengths(kind, n, targets)
    actions = H.random_actions(rng, T, B, len(gs), 0.1)
    coins = rng.integers(0, 2, size=(T, B))
    okw = dict(kw)
    pkw = {k: v for k, v in kw.items() if k in ("max_rotations",)}
    if kind != H.PAULI:
        okw["add_inverts"] = inverts
        pkw["add_inverts"] = inverts
    for b in range(B):
        a = orc.OracleEnv(kind, n, gs, add_perms=False, **okw)
        p = pyref.make(kind, n, gs, **pkw)
        st = targets[b, : lens[b]].tolist()
        a.set_state(st); p.set_state(st)
        assert a.observe() == p.observe()
        for t in range(T):
            c = bool(coins[t, b]) if (inverts and kind != H.PAULI) else None
            a.step(int(actions[t, b]), coin=c); p.step(int(actions[t, b]), coin=c)
            assert a.observe() == p.observe(), (name, b, t)
            assert bits(a.reward()) == bits(p.reward()), (name, b, t)
            assert a.is_final() == p.is_final() and a.masks() == p.masks()
            assert a.counts() == list(p.counts)
        assert a.solution() == p.solution()
        assert a.raw_state().tolist() == list(p.raw_state())


# ------------------------------------------------------------------ invariants
@pytest.mark.parametrize("name", ["C2_lf8_line", "C3_clifford8_full", "C1_perm_grid3", "clifford5_allgates"])
def test_solution_replays_to_identity_with_inverts(name):
    """solution() = solution ++ reverse(solution_inv) (clifford.rs:376-381): applying it to the target must give the
    identity whenever the episode ended solved, even with random inversions in between."""
    kind, n, gs, kw = H.config_table()[name]
    rng = np.random.Generator(np.random.PCG64(3))
    found = 0
    for trial in range(300):
        env = orc.OracleEnv(kind, n, gs, add_inverts=True, add_perms=False)
        env.difficulty = 3
        env.reset(seed=trial)
        start = env.raw_state().astype(np.int64).tolist()
        env.set_state(start)
        for t in range(12):
            if env.is_final():
                break
            env.step(int(rng.integers(0, len(gs))), coin=bool(rng.integers(0, 2)))
        if env.success() and len(env.solution()) > 0:
            found += 1
            rep = orc.OracleEnv(kind, n, gs, add_inverts=False, add_perms=False)
            rep.set_state(start)
            for a in env.solution():
                rep.step(a)
            assert rep.success(), (name, trial, env.solution())
    assert found > 0


def test_layer_sets_are_prefixes():
    """n_layers == max(last_gates)+1 (the GPU engine relies on this instead of hash sets)."""
    kind, n, gs, kw = H.config_table()["clifford5_allgates"]
    rng = np.random.Generator(np.random.PCG64(9))
    m = pyref.Metrics(n)
    for _ in range(500):
        g, q = gs[int(rng.integers(0, len(gs)))]
        m.apply(g, q)
        assert m.layers == set(range(max(m.last) + 1)) and m.cx_layers == set(range(max(m.last_cx) + 1))


def test_philox_known_answer():
    # Philox4x32-10 known-answer test (Random123 kat_vectors): counter = key = 0 -> 6627e8d5 ...
    assert orc.philox_draw(0, 0, 0, 0) == 0x6627E8D5
