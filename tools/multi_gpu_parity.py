#!/usr/bin/env python
"""Hardware multi-GPU differential test (SURVEY.md §4 plan item 4, §8e): the same GLOBAL batch sharded 1 / 2 / 4 / 8 ways must give
identical per-env outputs, and one `num_searches = R` search sharded over the ranks must return the same key and action list.

    python tools/multi_gpu_parity.py --out gpurun_out/r2_parity                          # N = 1: writes the reference arrays
    python -m torch.distributed.run --nproc-per-node N ... tools/multi_gpu_parity.py --out gpurun_out/r2_parity    # compares against them

Everything is keyed by GLOBAL env / rollout id: targets and actions are generated for the whole batch from one seed and each rank takes
its contiguous slice; `reset` and the search pass `first_env_id` / `first_rollout_id` = rank * slice.  Compared, bit for bit, per env:
f32 reward bits, done, success of every step, the final packed observation, the final state records, every solution, the reset states;
per search: the winning key, the global rollout id, the action list — through BOTH cross-GPU reductions (qg_search_finish over the C
ABI's own ncclComm_t, and search.reduce_best over torch.distributed) which must agree with each other and with the 1-GPU run.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r2_parity")
    ap.add_argument("--envs", type=int, default=8192, help="GLOBAL batch")
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--rollouts", type=int, default=1000, help="GLOBAL num_searches")
    args = ap.parse_args()

    from qiskit_gym_b200 import BatchedEnv, engine
    from qiskit_gym_b200 import workloads as W
    from qiskit_gym_b200.search import BasicPolicy, RolloutSearch

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/qg_parity_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
        comm = engine.nccl_comm_create(local)
    Bg, T = args.envs, args.steps
    assert Bg % world == 0 and args.rollouts % world == 0
    Bl = Bg // world
    lo = rank * Bl

    def gather(t: torch.Tensor, dim: int):
        """local shard (env axis = dim) -> global array on rank 0 (numpy), None elsewhere"""
        t = t.contiguous()
        if world == 1:
            return t.cpu().numpy()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return torch.cat(parts, dim=dim).cpu().numpy() if rank == 0 else None

    arrays = {}
    for cname, (kind, n, gateset, kw) in W.baseline_configs().items():
        for inv in ((False, True) if kind != W.PAULI else (False,)):
            tag = f"{cname}{'_inv' if inv else ''}"
            pk = dict(kw)
            if kind != W.PAULI:
                pk["add_inverts"] = inv
            seed = 20261017 + len(arrays)
            targets = W.random_targets(kind, n, gateset, Bg, seed, scramble=40)          # the GLOBAL batch, identical on every rank
            rng = np.random.Generator(np.random.PCG64(seed))
            actions = W.random_actions(rng, T, Bg, len(gateset), 0.02)
            coins = rng.integers(0, 2, size=(T, Bg)).astype(np.uint8) if inv else None
            env = BatchedEnv(kind, n, gateset, Bl, device=local, max_depth=T + 8, add_perms=False, **pk)
            env.set_state(targets[lo:lo + Bl])
            rew = torch.zeros((T, Bl), dtype=torch.float32, device=dev); don = torch.zeros((T, Bl), dtype=torch.bool, device=dev); suc = torch.zeros_like(don)
            env.replay(torch.from_numpy(np.ascontiguousarray(actions[:, lo:lo + Bl])).to(dev),
                       coins=None if coins is None else torch.from_numpy(np.ascontiguousarray(coins[:, lo:lo + Bl])).to(dev), reward=rew, done=don, success=suc)
            packed_ok = not (kind == W.PERM and n > 64)
            fin = env.observe_bits(env.new_obs_bits()) if packed_ok else env.observe().reshape(Bl, -1)
            sols = env.solutions()
            cap = T + 24
            sol_arr = np.full((Bl, cap), -1, dtype=np.int64)
            for i, s_ in enumerate(sols):
                sol_arr[i, :len(s_)] = s_
            met = env.metrics()
            arrays[f"{tag}/reward_bits"] = gather(rew.view(torch.int32), 1)
            arrays[f"{tag}/done"] = gather(don.to(torch.uint8), 1)
            arrays[f"{tag}/success"] = gather(suc.to(torch.uint8), 1)
            arrays[f"{tag}/final_obs"] = gather(fin, 0)
            arrays[f"{tag}/metrics"] = gather(met, 0)
            arrays[f"{tag}/solutions"] = gather(torch.from_numpy(sol_arr).to(dev), 0)
            # Env::reset keyed by global env id
            env.difficulty = 9
            env.reset(seed=4242, first_env_id=lo)
            fin0 = env.observe_bits(env.new_obs_bits()) if packed_ok else env.observe().reshape(Bl, -1)
            arrays[f"{tag}/reset_obs"] = gather(fin0, 0)
            arrays[f"{tag}/reset_depth"] = gather(env.status()[3], 0)
            del env

    # ---- one search of R rollouts sharded over the ranks (BASELINE.json configs[4], rl/synthesis.py:112-126) ---------------------------
    kind, n, gateset, kw = W.baseline_configs()["C5_perm27_heavyhex"]
    R = args.rollouts
    Rl = R // world
    torch.manual_seed(0)
    pol = BasicPolicy([n, n], len(gateset), embedding_size=512, common_layers=(256,))
    rs = RolloutSearch(kind, n, gateset, pol, Rl, device=local, max_depth=128, add_inverts=False, policy_backend="persistent")
    rng = np.random.Generator(np.random.PCG64(99))
    searches = []
    for s_i in range(6):
        # targets 1..3 random SWAPs of the map away from the identity (a random-init policy finds those), then uniform random ones
        perm = list(range(n))
        if s_i < 4:
            for _ in range(1 + s_i % 3):
                a, b = gateset[int(rng.integers(len(gateset)))][1]
                perm[a], perm[b] = perm[b], perm[a]
        else:
            perm = rng.permutation(n).tolist()
        res_c = rs.solve(perm, deterministic=False, seed=7 + s_i, first_rollout_id=rank * Rl, comm=comm)          # C ABI: qg_search_finish(ncclComm_t)
        res_t = rs.solve(perm, deterministic=False, seed=7 + s_i, first_rollout_id=rank * Rl)                     # torch.distributed: reduce_best
        assert (res_c.key, res_c.actions) == (res_t.key, res_t.actions), f"search {s_i}: the two cross-GPU reductions disagree on rank {rank}"
        searches.append({"target": perm, "key": int(res_c.key), "rollout_id": int(res_c.rollout_id), "success": bool(res_c.success), "actions": res_c.actions,
                         "iterations": int(res_c.iterations)})
    if world > 1:
        # every rank ended with the same winner
        mine = torch.tensor([s_["key"] for s_ in searches], dtype=torch.int64, device=dev)
        allk = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        assert all(torch.equal(allk[0], k) for k in allk), "ranks disagree on the winning keys"

    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        digest = {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() for k, v in arrays.items()}
        report = {"world": world, "global_envs": Bg, "steps": T, "rollouts": R, "arrays": len(arrays), "searches": searches}
        ref_npz, ref_json = args.out + "_ref.npz", args.out + "_ref.json"
        if world == 1:
            np.savez_compressed(ref_npz, **{k.replace("/", "__"): v for k, v in arrays.items()})
            json.dump({"digest": digest, "searches": searches}, open(ref_json, "w"))
            report["wrote_reference"] = True
            report["solved_searches"] = sum(s_["success"] for s_ in searches)
        else:
            ref = json.load(open(ref_json))
            bad = [k for k in digest if digest[k] != ref["digest"].get(k)]
            detail = {}
            if bad:
                z = np.load(ref_npz)
                for k in bad[:6]:
                    a, b = z[k.replace("/", "__")], arrays[k]
                    detail[k] = {"shape_ref": list(a.shape), "shape": list(b.shape), "first_diff": np.argwhere(a != b)[:4].tolist() if a.shape == b.shape else None}
            s_bad = [i for i, (a, b) in enumerate(zip(ref["searches"], searches)) if (a["key"], a["actions"], a["rollout_id"]) != (b["key"], b["actions"], b["rollout_id"])]
            report.update({"arrays_equal_to_1gpu": len(digest) - len(bad), "arrays_differ": bad, "detail": detail,
                           "searches_equal_to_1gpu": len(searches) - len(s_bad), "searches_differ": s_bad,
                           "match": not bad and not s_bad})
        with open(f"{args.out}_n{world}.json", "w") as f:
            json.dump(report, f)
        print(json.dumps({k: v for k, v in report.items() if k != "searches"}))
    if world > 1:
        engine.nccl_comm_destroy(comm)
        dist.destroy_process_group()
    if rank == 0 and world > 1 and not report["match"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
