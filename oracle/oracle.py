"""ctypes binding of the CPU oracle (oracle/libqg_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (qiskit_gym_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from qiskit_gym_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqg_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("qg_oracle_c.cpp", "qg_oracle.hpp", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "qg_engine.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.qgo_last_error.restype = C.c_char_p
        L.qgo_create.restype = C.c_void_p
        L.qgo_create.argtypes = [C.POINTER(_abi.QgConfig)]
        L.qgo_clone.restype = C.c_void_p
        L.qgo_clone.argtypes = [C.c_void_p]
        L.qgo_destroy.argtypes = [C.c_void_p]
        for name in ("qgo_num_actions", "qgo_get_difficulty", "qgo_is_final", "qgo_success"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int
        L.qgo_obs_shape.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.qgo_set_difficulty.argtypes = [C.c_void_p, C.c_int]
        L.qgo_set_state.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64]
        L.qgo_reset_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.qgo_step.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.qgo_observe.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_uint32]
        L.qgo_masks.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.qgo_reward.argtypes = [C.c_void_p]
        L.qgo_reward.restype = C.c_float
        L.qgo_depth.argtypes = [C.c_void_p]
        L.qgo_depth.restype = C.c_int64
        L.qgo_solution.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        L.qgo_raw_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int]
        L.qgo_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.qgo_twists.restype = C.c_int64
        L.qgo_twists.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.qgo_run_batch.argtypes = [C.POINTER(_abi.QgConfig), C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_void_p]
        L.qgo_bench.restype = C.c_double
        L.qgo_bench.argtypes = [C.POINTER(_abi.QgConfig), C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.qgo_digest.argtypes = [C.POINTER(_abi.QgConfig), C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.qgo_philox_draw.restype = C.c_uint32
        L.qgo_philox_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.qgo_gate_kind_from_name.argtypes = [C.c_char_p, C.c_int32]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_config(env_kind, num_qubits, gateset, difficulty=1, depth_slope=2, max_depth=128, **kw):
    gates = _abi.parse_gateset(gateset, lib().qgo_gate_kind_from_name)
    return _abi.make_config(env_kind, num_qubits, difficulty, gates, len(list(gateset)), depth_slope, max_depth, **kw)


class OracleEnv:
    """Single-env oracle object with the reference's raw-env method names."""

    def __init__(self, env_kind, num_qubits, gateset, difficulty=1, depth_slope=2, max_depth=128, **kw):
        self.cfg = make_config(env_kind, num_qubits, list(gateset), difficulty, depth_slope, max_depth, **kw)
        self._h = lib().qgo_create(C.byref(self.cfg))
        if not self._h:
            raise ValueError(lib().qgo_last_error().decode())
        shp = (C.c_int32 * 2)()
        lib().qgo_obs_shape(self._h, shp)
        self._shape = (shp[0], shp[1])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qgo_destroy(self._h)
            self._h = None

    def clone(self):
        """`Clone` of the env (the reference clones one env per rollout / tree node)."""
        o = object.__new__(OracleEnv)
        o.cfg, o._shape = self.cfg, self._shape
        o._h = lib().qgo_clone(self._h)
        return o

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(lib().qgo_last_error().decode())
        return rc

    def obs_shape(self):
        return list(self._shape)

    def num_actions(self):
        return lib().qgo_num_actions(self._h)

    @property
    def difficulty(self):
        return lib().qgo_get_difficulty(self._h)

    @difficulty.setter
    def difficulty(self, d):
        lib().qgo_set_difficulty(self._h, int(d))

    def set_state(self, state):
        arr = _abi.as_i64_array(list(state))
        self._chk(lib().qgo_set_state(self._h, arr, len(state)))

    def reset(self, seed=0, env_id=0):
        self._chk(lib().qgo_reset_philox(self._h, seed, env_id))

    def step(self, action, coin=None):
        self._chk(lib().qgo_step(self._h, int(action), -1 if coin is None else int(bool(coin))))

    def observe(self, perm_raw=None):
        cap = self._shape[0] * self._shape[1]
        out = (C.c_int64 * max(cap, 1))()
        n = self._chk(lib().qgo_observe(self._h, out, cap, 0 if perm_raw is None else 1, 0 if perm_raw is None else int(perm_raw)))
        return [out[i] for i in range(n)]

    def masks(self):
        out = (C.c_uint8 * max(self.num_actions(), 1))()
        n = lib().qgo_masks(self._h, out)
        return [bool(out[i]) for i in range(n)]

    def reward(self):
        return float(lib().qgo_reward(self._h))

    def is_final(self):
        return bool(lib().qgo_is_final(self._h))

    def success(self):
        return bool(lib().qgo_success(self._h))

    def depth(self):
        return int(lib().qgo_depth(self._h))

    def solution(self):
        cap = 1 << 16
        out = (C.c_int64 * cap)()
        n = lib().qgo_solution(self._h, out, cap)
        return [out[i] for i in range(n)]

    def raw_state(self):
        cap = 1 << 16
        out = (C.c_uint8 * cap)()
        n = lib().qgo_raw_state(self._h, out, cap)
        return np.frombuffer(out, dtype=np.uint8, count=n).copy()

    def counts(self):
        out = (C.c_int64 * 4)()
        lib().qgo_counts(self._h, out)
        return [out[i] for i in range(4)]

    def twists(self, internal=False):
        ol, al = C.c_int64(), C.c_int64()
        cnt = lib().qgo_twists(self._h, int(internal), None, 0, None, 0, C.byref(ol), C.byref(al))
        obs = np.zeros((cnt, ol.value), dtype=np.int64)
        act = np.zeros((cnt, al.value), dtype=np.int64)
        lib().qgo_twists(self._h, int(internal), _ptr(obs), obs.size, _ptr(act), act.size, C.byref(ol), C.byref(al))
        return obs.tolist(), act.tolist()


def pack_targets(targets):
    """list of int lists -> (int64[B, stride], lens[B])"""
    B = len(targets)
    stride = max(len(t) for t in targets)
    arr = np.zeros((B, stride), dtype=np.int64)
    lens = np.zeros(B, dtype=np.int64)
    for i, t in enumerate(targets):
        arr[i, : len(t)] = t
        lens[i] = len(t)
    return arr, lens


def run_batch(cfg, targets, lens, actions, coins=None, perm_raw=None, want_obs=True, state_cap=8192, sol_cap=1024):
    """Differential driver: see qgo_run_batch.  targets int64[B,stride]; actions int32[T,B]."""
    L = lib()
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    T, B = actions.shape
    h = L.qgo_create(C.byref(cfg))
    shp = (C.c_int32 * 2)()
    L.qgo_obs_shape(h, shp)
    L.qgo_destroy(h)
    OBS = shp[0] * shp[1]
    if coins is not None:
        coins = np.ascontiguousarray(coins, dtype=np.uint8)
    if perm_raw is not None:
        perm_raw = np.ascontiguousarray(perm_raw, dtype=np.uint32)
    out = {
        "obs0": np.zeros((B, OBS), np.uint8) if want_obs else None,
        "obs": np.zeros((T, B, OBS), np.uint8) if want_obs else None,
        "reward": np.zeros((T, B), np.float32),
        "done": np.zeros((T, B), np.uint8),
        "success": np.zeros((T, B), np.uint8),
        "counts": np.zeros((T, B, 4), np.int64),
        "depth": np.zeros((T, B), np.int64),
        "final_state": np.zeros((B, state_cap), np.uint8),
        "final_state_len": np.zeros(B, np.int64),
        "solutions": np.zeros((B, sol_cap), np.int64),
        "sol_len": np.zeros(B, np.int64),
    }
    rc = L.qgo_run_batch(C.byref(cfg), B, _ptr(targets), targets.shape[1], _ptr(lens), T, _ptr(actions), _ptr(coins), _ptr(perm_raw),
                         _ptr(out["obs0"]), _ptr(out["obs"]), _ptr(out["reward"]), _ptr(out["done"]), _ptr(out["success"]),
                         _ptr(out["counts"]), _ptr(out["depth"]),
                         _ptr(out["final_state"]), state_cap, _ptr(out["final_state_len"]),
                         _ptr(out["solutions"]), sol_cap, _ptr(out["sol_len"]))
    if rc < 0:
        raise RuntimeError(L.qgo_last_error().decode())
    return out


def bench(cfg, targets, lens, actions, coins=None, threads=1):
    """CPU baseline (see qgo_bench): returns (seconds, checksum)."""
    L = lib()
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    T, B = actions.shape
    if coins is not None:
        coins = np.ascontiguousarray(coins, dtype=np.uint8)
    cs = C.c_uint64()
    sec = L.qgo_bench(C.byref(cfg), B, _ptr(targets), targets.shape[1], _ptr(lens), T, _ptr(actions), _ptr(coins), threads, C.byref(cs))
    if sec < 0:
        raise RuntimeError(L.qgo_last_error().decode())
    return sec, cs.value


def digest(cfg, targets, lens, actions, wobs, wmask, coins=None, threads=1):
    """Per-env uint64 digest over every output of every step (see qgo_digest).  wobs uint64[OBS], wmask uint64[A], all < 2^31."""
    L = lib()
    targets = np.ascontiguousarray(targets, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    wobs = np.ascontiguousarray(wobs, dtype=np.uint64)
    wmask = np.ascontiguousarray(wmask, dtype=np.uint64)
    T, B = actions.shape
    if coins is not None:
        coins = np.ascontiguousarray(coins, dtype=np.uint8)
    out = np.zeros(B, np.uint64)
    rc = L.qgo_digest(C.byref(cfg), B, _ptr(targets), targets.shape[1], _ptr(lens), T, _ptr(actions), _ptr(coins), threads, _ptr(wobs), _ptr(wmask), _ptr(out))
    if rc < 0:
        raise RuntimeError(L.qgo_last_error().decode())
    return out


def philox_draw(seed, env, idx, stream):
    return int(lib().qgo_philox_draw(seed, env, idx, stream))
