"""The reference's own Python layer, executed UNMODIFIED, over this repo's raw-env classes.

`tests/golden/reference_python_kats.json` is what `/root/reference/src/qiskit_gym/envs/{adapters,synthesis}.py` produce when their one
native import (`qiskit_gym.qiskit_gym_rs`) is bound to the CPU oracle through `qiskit_gym_b200.reference_shim` (generator:
tests/golden/make_reference_python_golden.py).  Here:
  * not-gpu, reference tree present: the generator is re-run live and must reproduce the committed fixture (the reference files are
    imported from their own path, nothing is re-typed);
  * not-gpu: this repo's Qiskit-free wire helpers agree with what the reference's functions returned;
  * gpu: the same episodes on the CUDA engine's raw-env classes (`qiskit_gym_b200.envs`, through the C ABI) give the same spaces,
    dense observations, f32 reward bits, terminated flags, solutions and decoded PauliNetwork rotations.  Where the reference tree is
    reachable the episodes go through the reference's own `GymWrapper`; on the GPU box (no /root/reference there) they go through the
    raw-env protocol the wrapper documents (adapters.py:22-33: obs_shape, observe, reward, is_final, num_actions, reset, step).
"""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference_python_kats.json")
REF_SRC = os.environ.get("QISKIT_GYM_SRC", "/root/reference/src")
HAVE_REF = os.path.isfile(os.path.join(REF_SRC, "qiskit_gym", "envs", "adapters.py"))


@pytest.fixture(scope="module")
def kats():
    return json.load(open(FIXTURE))


def f32_bits(x) -> int:
    return int(np.float32(x).view(np.uint32))


@pytest.mark.skipif(not HAVE_REF, reason="the reference tree is not present on this machine")
def test_fixture_is_what_the_unmodified_reference_python_produces(kats):
    from tests.golden import make_reference_python_golden as gen
    live = json.loads(json.dumps(gen.generate()))
    assert live == kats


def test_shim_refuses_to_invent_the_reference(tmp_path):
    from qiskit_gym_b200 import reference_shim as shim
    import types
    backend = types.ModuleType("fake_rs")
    with pytest.raises(ImportError):
        shim.install(str(tmp_path), backend=backend)          # lacks the four class names
    for nm in ("PermutationEnv", "LinearFunctionEnv", "CliffordEnv", "PauliNetworkEnv"):
        setattr(backend, nm, object)
    old = os.environ.pop("QISKIT_GYM_SRC", None)
    try:
        import importlib.util
        if importlib.util.find_spec("qiskit_gym") is None:
            with pytest.raises(ImportError):
                shim.install(str(tmp_path), backend=backend)  # no qiskit_gym package there
    finally:
        if old is not None:
            os.environ["QISKIT_GYM_SRC"] = old


def test_wire_helpers_agree_with_reference_outputs(kats):
    from qiskit_gym_b200 import wire
    from qiskit_gym_b200 import workloads as W
    assert wire.ONE_Q_GATES == kats["constants"]["ONE_Q_GATES"] and wire.TWO_Q_GATES == kats["constants"]["TWO_Q_GATES"]
    assert wire.ROTATION_MARKER == kats["constants"]["ROTATION_MARKER"]
    for c in kats["cases"]:
        allowed = kats["constants"]["allowed_gates"][c["cls"]]
        n, gs = W.gateset_from_coupling_map([tuple(e) for e in c["edges"]], c["basis_gates"] or allowed)
        assert n == c["config"]["num_qubits"]
        assert [[g, list(q)] for g, q in gs] == c["config"]["gateset"]
        if "decoded_solution" in c:
            assert [list(t) for t in wire.decode_pauli_solution(c["solution"])] == c["decoded_solution"]
    for e in kats["perm_get_state"]:
        assert wire.permutation_state(e["pattern"]).tolist() == e["state"]
    dw = kats["decode_words"]
    assert [list(t) for t in wire.decode_pauli_solution(dw["words"])] == dw["decoded"]


def _raw_class(c):
    from qiskit_gym_b200 import envs
    return {"PermutationGym": envs.PermutationEnv, "LinearFunctionGym": envs.LinearFunctionEnv, "CliffordGym": envs.CliffordEnv,
            "PauliGym": envs.PauliNetworkEnv}[c["cls"]]


class _Protocol:
    """Drives a raw env the way the reference's wrapper does (adapters.py:50-72), for machines without the reference tree."""

    def __init__(self, raw):
        self.raw = raw
        self.shape = tuple(raw.obs_shape())

    def full_obs(self):
        full = np.zeros(int(np.prod(self.shape)), dtype=np.int8)
        full[self.raw.observe()] = 1
        return full.reshape(self.shape)

    def step(self, a):
        assert not bool(self.raw.is_final())
        self.raw.step(int(a))
        return self.full_obs(), float(self.raw.reward()), bool(self.raw.is_final())


@pytest.mark.gpu
def test_engine_replays_the_reference_wrapper_traces(kats):
    """Every recorded episode, on the CUDA engine.  With the reference tree present the engine sits under the reference's own classes."""
    syn = None
    made = []
    if HAVE_REF:
        from qiskit_gym_b200 import reference_shim as shim
        from tests import ref_stubs
        made = ref_stubs.install_third_party_stubs()
        shim.uninstall()
        shim.install(REF_SRC)
        import qiskit_gym.envs.synthesis as syn
    try:
        for c in kats["cases"]:
            cfg = c["config"]
            if syn is not None:
                env = getattr(syn, c["cls"]).from_coupling_map([tuple(e) for e in c["edges"]], basis_gates=None if c["basis_gates"] is None else tuple(c["basis_gates"]), **c["kwargs"])
                assert json.loads(json.dumps(env.to_json())) == cfg
                assert list(env.observation_space.shape) == c["observation_space_shape"] and int(env.action_space.n) == c["action_space_n"]
                raw = env._raw_env
                full_obs, step = env._full_obs, (lambda a, env=env: env.step(a)[:3])
            else:
                kw = {k: v for k, v in cfg.items() if k not in ("num_qubits", "difficulty", "gateset", "depth_slope", "max_depth")}
                if c["cls"] == "PauliGym":
                    raw = _raw_class(c)(cfg["num_qubits"], cfg["difficulty"], cfg["gateset"], cfg["depth_slope"], cfg["max_depth"], kw.pop("max_rotations"), **kw)
                else:
                    raw = _raw_class(c)(cfg["num_qubits"], cfg["difficulty"], cfg["gateset"], cfg["depth_slope"], cfg["max_depth"], **kw)
                drv = _Protocol(raw)
                full_obs, step = drv.full_obs, drv.step
                assert list(raw.obs_shape()) == c["observation_space_shape"] and raw.num_actions() == c["action_space_n"]
            raw.set_state(c["state"])
            assert np.flatnonzero(full_obs().reshape(-1)).tolist() == c["obs0"], c["cls"]
            for t, rec in enumerate(c["trace"]):
                obs, reward, terminated = step(rec["a"])
                assert obs.dtype == np.int8
                assert np.flatnonzero(obs.reshape(-1)).tolist() == rec["obs"], (c["cls"], t)
                assert f32_bits(reward) == rec["reward_bits"], (c["cls"], t)
                assert terminated == rec["terminated"], (c["cls"], t)
            assert bool(raw.success()) == c["success"]
            assert [int(v) for v in raw.solution()] == c["solution"]
            if "final_assert" in c:
                assert raw.is_final()
            if "decoded_solution" in c:
                from qiskit_gym_b200 import wire
                dec = syn.decode_pauli_solution if syn is not None else wire.decode_pauli_solution
                assert [list(x) for x in dec(raw.solution())] == c["decoded_solution"]
            raw.difficulty = 3
            assert raw.difficulty == c["difficulty_after_set"]
            raw.reset()
            assert full_obs().shape == tuple(c["observation_space_shape"])
    finally:
        if HAVE_REF:
            from qiskit_gym_b200 import reference_shim as shim
            from tests import ref_stubs
            shim.uninstall()
            ref_stubs.remove_stubs(made)
