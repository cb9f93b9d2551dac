"""Host-only entry points of the C ABI (no GPU needed): config validation, obs_shape, twists."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import check, lib


def make_config(env_kind, num_qubits, gateset, difficulty=1, depth_slope=2, max_depth=128, **kw):
    gateset = list(gateset)
    gates = _abi.parse_gateset(gateset, lib().qg_gate_kind_from_name)
    return _abi.make_config(env_kind, num_qubits, difficulty, gates, len(gateset), depth_slope, max_depth, **kw)


def validate(cfg) -> None:
    check(lib().qg_config_validate(C.byref(cfg)))


def obs_shape(cfg):
    shp = (C.c_int32 * 2)()
    check(lib().qg_config_obs_shape(C.byref(cfg), shp))
    return [int(shp[0]), int(shp[1])]


def twists(cfg):
    """Env::twists for a config: (obs_perms, act_perms) as lists of lists (symmetry.rs:297-361)."""
    L = lib()
    t = C.c_void_p()
    check(L.qg_twists_create(C.byref(cfg), C.byref(t)))
    try:
        cnt, ol, al = L.qg_twists_count(t), L.qg_twists_obs_len(t), L.qg_twists_act_len(t)
        obs = np.zeros((cnt, ol), dtype=np.int64)
        act = np.zeros((cnt, al), dtype=np.int64)
        check(L.qg_twists_copy(t, obs.ctypes.data_as(C.c_void_p), act.ctypes.data_as(C.c_void_p)))
    finally:
        L.qg_twists_destroy(t)
    return obs.tolist(), act.tolist()
