"""GPU tests of the C-ABI entry points around the fused step (csrc/qg_extras.cu): bulk solution read-out, the packed host wire
format (uint8 actions in, f32 reward + 2-bit-per-env-step flag planes out), NUMA-local pinned buffers, the DLPack view of the
observation ring and the single-GPU end of qg_search_finish — each against the oracle or against the plain entry point it narrows."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _setup(name, B, T, seed, add_inverts=False):
    kind, n, gateset, kw = H.config_table()[name]
    pk = dict(kw)
    if kind != H.PAULI:
        pk["add_inverts"] = add_inverts
    rng = np.random.Generator(np.random.PCG64(seed))
    cfg = H.make_cfg(kind, n, gateset, add_perms=False, **pk)
    tarr = H.random_targets(kind, n, gateset, B, seed, scramble=12, num_rotations=kw.get("max_rotations", 5) + 1, vary_rotations=True)
    lens = H.payload_lengths(kind, n, tarr)
    actions = H.random_actions(rng, T, B, len(gateset), 0.03)
    coins = rng.integers(0, 2, size=(T, B)).astype(np.uint8) if (add_inverts and kind != H.PAULI) else None
    return kind, n, gateset, pk, cfg, tarr, lens, actions, coins


@pytest.mark.parametrize("name,inv", [("C2_lf8_line", True), ("C3_clifford8_full", True), ("C1_perm_grid3", True), ("pauli6_line", False), ("lf11_line", False)])
def test_bulk_solutions_equal_oracle_and_single_reads(name, inv):
    from qiskit_gym_b200 import BatchedEnv
    B, T = 200, 30
    kind, n, gateset, pk, cfg, tarr, lens, actions, coins = _setup(name, B, T, 77, inv)
    ref = orc.run_batch(cfg, tarr, lens, actions, coins=coins, want_obs=False)
    env = BatchedEnv(kind, n, gateset, B, add_perms=False, **pk)
    env.set_state(tarr)
    env.replay(torch.from_numpy(actions).to(env.device), coins=None if coins is None else torch.from_numpy(coins).to(env.device))
    sols = env.solutions()
    assert len(sols) == B
    for b in range(B):
        assert sols[b] == ref["solutions"][b, : int(ref["sol_len"][b])].tolist(), (name, b)
    for b in (0, 31, 32, B - 1):
        assert env.solution(b) == sols[b]
    assert env.solutions(first=37, count=5) == sols[37:42]
    with pytest.raises(ValueError):
        env.solutions(cap=2)


@pytest.mark.parametrize("tile", [16, 32])
@pytest.mark.parametrize("name,inv,B,T", [("C3_clifford8_full", False, 333, 70), ("C2_lf8_line", True, 64, 33), ("C4_pauli10_line", False, 97, 40),
                                          ("C5_perm27_heavyhex", False, 1, 5),
                                          # >= 1 MB of actions: the host form stages them through the copy engine in flagged chunks (ragged rows)
                                          ("C3_clifford8_full", False, 16391, 70), ("C2_lf8_line", True, 32773, 40)])
def test_packed_wire_format_equals_plain_replay(name, inv, B, T, tile):
    """uint8 actions / transposed flag bit planes carry exactly what int32 actions / uint8 flags carry (device and pinned-host forms)."""
    from qiskit_gym_b200 import BatchedEnv
    kind, n, gateset, pk, cfg, tarr, lens, actions, coins = _setup(name, B, T, 5, inv)
    actions = np.minimum(actions, 255).astype(np.int32)          # (an invalid action stays invalid: num_actions <= 104 here)
    env = BatchedEnv(kind, n, gateset, B, add_perms=False, tile_envs=tile, **pk)
    dev = env.device
    env.set_state(tarr)
    env.snapshot()
    a32 = torch.from_numpy(actions).to(dev)
    c = None if coins is None else torch.from_numpy(coins).to(dev)
    rew = torch.zeros((T, B), dtype=torch.float32, device=dev); don = torch.zeros((T, B), dtype=torch.bool, device=dev); suc = torch.zeros_like(don)
    obs = torch.zeros((2, B) + tuple(env.obs_shape()), dtype=torch.float32, device=dev)
    env.replay(a32, coins=c, obs=obs, reward=rew, done=don, success=suc)
    final_plain = [env.get_state(b) for b in (0, B // 2, B - 1)]
    # device-resident packed form
    env.restore()
    a8 = torch.from_numpy(actions.astype(np.uint8)).to(dev)
    tiles = env.flag_words()
    db = torch.full((tiles, T), -1, dtype=torch.int32, device=dev); sb = torch.full((tiles, T), -1, dtype=torch.int32, device=dev)
    rew2 = torch.zeros_like(rew); obs2 = torch.zeros_like(obs)
    env.replay_packed(a8, done_bits=db, success_bits=sb, reward=rew2, coins=c, obs=obs2)
    assert torch.equal(rew2.view(torch.int32), rew.view(torch.int32)) and torch.equal(obs2, obs)
    d = env.unpack_flag_bits(db.cpu().numpy().view(np.uint32), B); s = env.unpack_flag_bits(sb.cpu().numpy().view(np.uint32), B)
    assert np.array_equal(d, don.cpu().numpy().astype(np.uint8)) and np.array_equal(s, suc.cpu().numpy().astype(np.uint8))
    assert [env.get_state(b).tolist() for b in (0, B // 2, B - 1)] == [f.tolist() for f in final_plain]
    # pinned host form (NUMA-local buffers from the engine)
    env.restore()
    h_a = env.host_buffer((T, B), np.uint8); h_r = env.host_buffer((T, B), np.float32)
    h_d = env.host_buffer((tiles, T), np.uint32); h_s = env.host_buffer((tiles, T), np.uint32)
    h_c = None
    if coins is not None:
        h_c = env.host_buffer((T, B), np.uint8); h_c[:] = coins
    h_a[:] = actions.astype(np.uint8)
    env.replay_host_packed(h_a, h_d, h_s, reward=h_r, coins=h_c, obs=obs2)
    assert np.array_equal(h_r.view(np.uint32), rew.cpu().numpy().view(np.uint32))
    assert np.array_equal(env.unpack_flag_bits(h_d, B), d) and np.array_equal(env.unpack_flag_bits(h_s, B), s)
    # rewards kept on the device, flags only to the host
    env.restore()
    rew3 = torch.zeros_like(rew); h_d[:] = 0
    env.replay_host_packed(h_a, h_d, None, reward_dev=rew3, coins=h_c)
    assert torch.equal(rew3.view(torch.int32), rew.view(torch.int32)) and np.array_equal(env.unpack_flag_bits(h_d, B), d)
    # two episodes in flight (qg_replay_host_packed_async), each with its own output buffers
    h_r2 = env.host_buffer((T, B), np.float32); h_d2 = env.host_buffer((tiles, T), np.uint32)
    h_r[:] = 0; h_d[:] = 0; h_r2[:] = 0; h_d2[:] = 0
    env.restore(); env.replay_host_packed(h_a, h_d, None, reward=h_r, coins=h_c, sync=False)
    env.restore(); env.replay_host_packed(h_a, h_d2, None, reward=h_r2, coins=h_c, sync=False)
    torch.cuda.synchronize()
    for hr, hd in ((h_r, h_d), (h_r2, h_d2)):
        assert np.array_equal(hr.view(np.uint32), rew.cpu().numpy().view(np.uint32)) and np.array_equal(env.unpack_flag_bits(hd, B), d)
    # pageable memory is refused, loudly
    with pytest.raises(ValueError):
        env.replay_host_packed(np.zeros((T, B), np.uint8), np.zeros((tiles, T), np.uint32))


def test_dlpack_observation_ring_is_zero_copy():
    from qiskit_gym_b200 import BatchedEnv
    B, T = 50, 6
    kind, n, gateset, pk, cfg, tarr, lens, actions, coins = _setup("C3_clifford8_full", B, T, 9)
    ref = orc.run_batch(cfg, tarr, lens, actions)
    env = BatchedEnv(kind, n, gateset, B, add_perms=False, **pk)
    env.set_state(tarr)
    view = env.dlpack_obs()
    assert view.is_cuda and view.dtype == torch.float32 and tuple(view.shape) == (B,) + tuple(env.obs_shape())
    assert env.dlpack_obs().data_ptr() == view.data_ptr()               # the same engine-owned memory every time
    for t in range(T):
        env.step(torch.from_numpy(actions[t]).to(env.device), obs=view, mask=False)
        assert np.array_equal(view.reshape(B, -1).cpu().numpy().astype(np.uint8), ref["obs"][t])
    with pytest.raises(ValueError):
        env.dlpack_obs(ring=3)                                          # one ring per engine


def test_search_finish_single_gpu_equals_best_plus_solution():
    from qiskit_gym_b200 import BatchedEnv
    kind, n, gateset, kw = H.config_table()["C1_perm_grid3"]
    B = 96
    env = BatchedEnv(kind, n, gateset, B, max_depth=24, add_inverts=False, add_perms=False)
    target = [1, 0, 2, 3, 4, 5, 6, 8, 7]
    env.set_state(target)
    env.search_begin(3, 1000)
    w = torch.ones((B, len(gateset)), dtype=torch.float32, device=env.device)
    for _ in range(24):
        env.search_step(w, deterministic=False, obs=False)
    key, idx = env.search_best()
    fk, ok, rid, owner, acts = env.search_finish()
    assert fk == key and rid == 1000 + idx and owner == 0
    assert ok == bool((key >> 62) & 1)
    if ok:
        assert acts == env.solution(idx)
    # no successful rollout -> no actions
    env.set_state([8, 7, 6, 5, 4, 3, 2, 1, 0])
    env.search_begin(3, 0)
    env.search_step(w, deterministic=False, obs=False)
    fk, ok, rid, owner, acts = env.search_finish()
    assert not ok and acts is None and fk != 0
