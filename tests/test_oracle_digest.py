"""The whole-batch digest used by tests/test_full_size.py (oracle/qg_oracle_c.cpp: qgo_digest): its tensor-side twin, fed with the oracle's own
per-step outputs, gives the same uint64 per environment, and the digest does not depend on the thread partition.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H
from tests.test_full_size import _device_digest


@pytest.mark.parametrize("name", ["C1_perm_grid3", "C2_lf8_line", "C3_clifford8_full", "C4_pauli10_line"])
def test_digest_twin_matches_oracle_digest(name):
    kind, n, gateset, kw = H.config_table()[name]
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = False
    cfg = H.make_cfg(kind, n, gateset, add_perms=False, **pk)
    B, T, A = 24, 16, len(gateset)
    tarr = H.random_targets(kind, n, gateset, B, 3, scramble=16, num_rotations=kw.get("max_rotations", 5))
    lens = H.payload_lengths(kind, n, tarr)
    rng = np.random.Generator(np.random.PCG64(1))
    acts = H.random_actions(rng, T, B, A, 0.05)
    ref = orc.run_batch(cfg, tarr, lens, acts)
    wobs = rng.integers(1, 2 ** 31, size=ref["obs"].shape[-1], dtype=np.uint64)
    wmask = rng.integers(1, 2 ** 31, size=A, dtype=np.uint64)
    d1 = orc.digest(cfg, tarr, lens, acts, wobs, wmask, threads=1)
    assert np.array_equal(d1, orc.digest(cfg, tarr, lens, acts, wobs, wmask, threads=5))
    masks = np.zeros((T, B, A), np.uint8)
    for b in range(B):
        e = orc.OracleEnv(kind, n, gateset, add_perms=False, **pk)
        e.set_state(tarr[b][: lens[b]])
        for t in range(T):
            e.step(int(acts[t, b]))
            masks[t, b] = np.array(e.masks(), dtype=np.uint8)
    got = _device_digest(torch.from_numpy(ref["obs"].astype(np.float32)), torch.from_numpy(masks.astype(bool)), torch.from_numpy(ref["reward"]),
                         torch.from_numpy(ref["done"].astype(bool)), torch.from_numpy(ref["success"].astype(bool)), wobs, wmask).numpy().view(np.uint64)
    assert np.array_equal(got, d1)
    # and it does see a single flipped observation entry
    ref["obs"][T // 2, 3, 0] ^= 1
    got2 = _device_digest(torch.from_numpy(ref["obs"].astype(np.float32)), torch.from_numpy(masks.astype(bool)), torch.from_numpy(ref["reward"]),
                          torch.from_numpy(ref["done"].astype(bool)), torch.from_numpy(ref["success"].astype(bool)), wobs, wmask).numpy().view(np.uint64)
    assert (got2 != d1).sum() == 1 and got2[3] != d1[3]
