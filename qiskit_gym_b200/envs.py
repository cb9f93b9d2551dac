"""Drop-in raw-env classes: the four pyo3 classes of `qiskit_gym.qiskit_gym_rs`
(rust/src/lib.rs:24-30) re-expressed over the batched CUDA engine with batch = 1.

Constructor signatures and defaults follow permutation.rs:266-299, linear_function.rs:373-406,
clifford.rs:390-423 and pauli.rs:728-775; the method set is the `PyBaseEnv` surface the reference's
Python code relies on (envs/adapters.py:22-33, rl/synthesis.py:97-109):
obs_shape, num_actions, observe (sparse indices), masks, step, reset, set_state, is_final, reward,
success, twists, solution, track_solution and the `difficulty` property — so
`qiskit_gym.envs.adapters.gym_adapter` wraps them unmodified.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _abi
from .engine import BatchedEnv


class _RawEnv:
    _KIND = -1

    def _build(self, num_qubits, difficulty, gateset, depth_slope, max_depth, **kw):
        self._track = True if kw.get("track_solution") is None else bool(kw.get("track_solution"))
        self._env = BatchedEnv(self._KIND, num_qubits, gateset, 1, difficulty=difficulty, depth_slope=depth_slope,
                               max_depth=max_depth, **kw)
        self._act = torch.zeros(1, dtype=torch.int32, device=self._env.device)
        self._coin = torch.zeros(1, dtype=torch.uint8, device=self._env.device)
        self._needs_coin = bool(self._env.cfg.add_inverts)
        self._perm_raw = torch.zeros(1, dtype=torch.int32, device=self._env.device)      # bit pattern of the uint32 draw

    # ---- Env trait surface -------------------------------------------------------------
    def obs_shape(self):
        return self._env.obs_shape()

    def num_actions(self):
        return self._env.num_actions()

    @property
    def difficulty(self):
        return self._env.difficulty

    @difficulty.setter
    def difficulty(self, value):
        self._env.difficulty = int(value)

    def set_state(self, state):
        self._env.set_state([int(x) for x in state])

    def reset(self, seed: int | None = None):
        """Env::reset.  The reference draws from an unseeded thread_rng; a fresh 64-bit seed is drawn
        from the OS unless one is given."""
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self._env.reset(seed, 0)

    def step(self, action, coin: bool | None = None):
        """Env::step.  `coin` injects the add_inverts coin flip (tests); default: drawn from the OS."""
        a = int(action)
        if a < 0:
            raise OverflowError("can't convert negative int to unsigned")
        self._act.fill_(min(a, 2**31 - 1))
        coins = None
        if self._needs_coin:
            if coin is None:
                coin = bool(os.urandom(1)[0] & 1)
            self._coin.fill_(1 if coin else 0)
            coins = self._coin
        self._env.step(self._act, coins=coins, obs=False, mask=False)

    def observe(self):
        perm_raw = None
        if self._KIND == _abi.ENV_PAULI_NETWORK and self._env.cfg.add_perms:
            # pauli.rs:656-660 draws a fresh qubit permutation from thread_rng on every observe(): the raw 32-bit draw comes from the OS
            self._perm_raw.fill_(int.from_bytes(os.urandom(4), "little") - (1 << 31))
            perm_raw = self._perm_raw
        obs = self._env.observe(perm_raw=perm_raw)
        return torch.nonzero(obs.reshape(-1), as_tuple=False).reshape(-1).tolist()

    def masks(self):
        return [bool(v) for v in self._env.masks().reshape(-1).tolist()]

    def is_final(self):
        return bool(self._env.status()[1].item())

    def reward(self):
        return float(self._env.status()[0].item())

    def success(self):
        return bool(self._env.status()[2].item())

    def twists(self):
        return self._env.twists()

    def track_solution(self):
        return self._track

    def solution(self):
        return self._env.solution(0)

    # ---- extras -----------------------------------------------------------------------
    def get_state(self):
        return self._env.get_state(0)

    def metrics(self):
        return [int(v) for v in self._env.metrics()[0].tolist()]


class PermutationEnv(_RawEnv):
    _KIND = _abi.ENV_PERMUTATION

    def __init__(self, num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=None,
                 add_inverts=None, add_perms=None, track_solution=None):
        self._build(num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=metrics_weights,
                    add_inverts=add_inverts, add_perms=add_perms, track_solution=track_solution)


class LinearFunctionEnv(_RawEnv):
    _KIND = _abi.ENV_LINEAR_FUNCTION

    def __init__(self, num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=None,
                 add_inverts=None, add_perms=None, track_solution=None):
        self._build(num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=metrics_weights,
                    add_inverts=add_inverts, add_perms=add_perms, track_solution=track_solution)


class CliffordEnv(_RawEnv):
    _KIND = _abi.ENV_CLIFFORD

    def __init__(self, num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=None,
                 add_inverts=None, add_perms=None, track_solution=None):
        self._build(num_qubits, difficulty, gateset, depth_slope, max_depth, metrics_weights=metrics_weights,
                    add_inverts=add_inverts, add_perms=add_perms, track_solution=track_solution)


class PauliNetworkEnv(_RawEnv):
    _KIND = _abi.ENV_PAULI_NETWORK

    def __init__(self, num_qubits, difficulty, gateset, depth_slope, max_depth, max_rotations,
                 pauli_diff_scale=None, num_qubits_decay=None, final_pauli_layers=None, metrics_weights=None,
                 add_perms=None, pauli_layer_reward=None, track_solution=None):
        self._build(num_qubits, difficulty, gateset, depth_slope, max_depth, max_rotations=max_rotations,
                    pauli_diff_scale=pauli_diff_scale, num_qubits_decay=num_qubits_decay,
                    final_pauli_layers=final_pauli_layers, metrics_weights=metrics_weights, add_perms=add_perms,
                    pauli_layer_reward=pauli_layer_reward, track_solution=track_solution)
