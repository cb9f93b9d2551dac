"""Plugs the CUDA engine in underneath the reference's own, unmodified Python layer.

The reference's Gymnasium wrappers and synthesis classes (src/qiskit_gym/envs/adapters.py:18-105,
src/qiskit_gym/envs/synthesis.py:66-528) reach the Rust core through exactly one import,
`from qiskit_gym import qiskit_gym_rs` (envs/synthesis.py:15), and use four names of it: `PermutationEnv`,
`LinearFunctionEnv`, `CliffordEnv`, `PauliNetworkEnv` (envs/synthesis.py:158, 223, 264, 308).  `qiskit_gym_b200.envs`
provides those four classes with the pyo3 constructor signatures over the C ABI, so the drop-in is to register it
under that module name BEFORE `qiskit_gym.envs` is imported:

    import qiskit_gym_b200.reference_shim as shim
    shim.install()                                   # qiskit_gym.qiskit_gym_rs -> qiskit_gym_b200.envs
    from qiskit_gym.envs import CliffordGym          # the reference's file, byte for byte
    env = CliffordGym.from_coupling_map(...)         # steps run on the GPU

Nothing of the reference's Python is restated here: this module only edits `sys.modules` / `sys.path`.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

RS_MODULE = "qiskit_gym.qiskit_gym_rs"


def _package_dir(src_dir: str | None) -> str | None:
    """Directory that contains the reference's `qiskit_gym/` package: `src_dir`, $QISKIT_GYM_SRC, or wherever an installed
    `qiskit_gym` distribution lives (located without importing it: its `__init__` is empty, but its compiled submodule is
    what is being replaced)."""
    for cand in (src_dir, os.environ.get("QISKIT_GYM_SRC")):
        if cand and os.path.isfile(os.path.join(cand, "qiskit_gym", "envs", "synthesis.py")):
            return os.path.abspath(cand)
    try:
        spec = importlib.util.find_spec("qiskit_gym")
    except (ImportError, ValueError):
        spec = None
    if spec is not None and spec.submodule_search_locations:
        return os.path.dirname(list(spec.submodule_search_locations)[0])
    return None


def install(src_dir: str | None = None, backend: types.ModuleType | None = None) -> types.ModuleType:
    """Registers `backend` (default: `qiskit_gym_b200.envs`, the engine-backed raw-env classes) as `qiskit_gym.qiskit_gym_rs`
    and makes the reference's `qiskit_gym` package importable from `src_dir` if it is not installed.  Returns the backend.
    Raises ImportError when the reference's Python package cannot be found: there is no re-typed copy to fall back to."""
    if backend is None:
        from . import envs as backend          # needs the CUDA library; fails loudly without it
    for name in ("PermutationEnv", "LinearFunctionEnv", "CliffordEnv", "PauliNetworkEnv"):
        if not hasattr(backend, name):
            raise ImportError(f"backend module lacks {name}")
    pkg_parent = _package_dir(src_dir)
    if pkg_parent is None:
        raise ImportError("the reference's qiskit_gym Python package was not found (pass src_dir= or set QISKIT_GYM_SRC)")
    if pkg_parent not in sys.path:
        sys.path.insert(0, pkg_parent)
    already = sys.modules.get("qiskit_gym.envs.synthesis")
    if already is not None and getattr(already, "qiskit_gym_rs", backend) is not backend:
        raise ImportError("qiskit_gym.envs was imported before reference_shim.install(): it is bound to another backend")
    sys.modules[RS_MODULE] = backend
    pkg = importlib.import_module("qiskit_gym")
    pkg.qiskit_gym_rs = backend
    return backend


def uninstall() -> None:
    """Drops the reference modules loaded through the shim (tests swap backends)."""
    for name in [m for m in sys.modules if m == "qiskit_gym" or m.startswith("qiskit_gym.")]:
        del sys.modules[name]


def synth_envs(src_dir: str | None = None, backend: types.ModuleType | None = None) -> dict:
    """`qiskit_gym.envs.synthesis.SYNTH_ENVS` (envs/synthesis.py:523-528) of the reference, bound to the engine."""
    install(src_dir, backend)
    return importlib.import_module("qiskit_gym.envs.synthesis").SYNTH_ENVS
