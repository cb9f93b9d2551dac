"""qiskit_gym_b200 — B200-native batched engine for qiskit-gym's synthesis environments.

Public surface
  BatchedEnv                         B environments, one fused CUDA launch per step (engine.py)
  PermutationEnv, LinearFunctionEnv,
  CliffordEnv, PauliNetworkEnv       drop-in raw-env classes of qiskit_gym.qiskit_gym_rs (envs.py)
"""
from . import _abi
from ._abi import ENV_CLIFFORD, ENV_LINEAR_FUNCTION, ENV_PAULI_NETWORK, ENV_PERMUTATION

__all__ = [
    "BatchedEnv", "PermutationEnv", "LinearFunctionEnv", "CliffordEnv", "PauliNetworkEnv",
    "ENV_PERMUTATION", "ENV_LINEAR_FUNCTION", "ENV_CLIFFORD", "ENV_PAULI_NETWORK",
]


def __getattr__(name):
    # torch (and the CUDA library) are only imported when an engine class is touched
    if name == "BatchedEnv":
        from .engine import BatchedEnv
        return BatchedEnv
    if name in ("PermutationEnv", "LinearFunctionEnv", "CliffordEnv", "PauliNetworkEnv"):
        from . import envs
        return getattr(envs, name)
    raise AttributeError(name)
