#!/usr/bin/env python
"""Development probe of the tensor-core policy kernel: correctness on one shape, then time per 65 536-row decision against the PyTorch
module (f32 / tf32 / bf16 cuBLAS).   timeout 120 python tools/tc_policy_probe.py [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from qiskit_gym_b200.policy import FusedPolicy, TensorCorePolicy, pack_obs_bits
from qiskit_gym_b200.search import BasicPolicy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
torch.manual_seed(0)
obs_shape, A = (16, 16), 72
pol = BasicPolicy(list(obs_shape), A, embedding_size=512, common_layers=(256,)).to(dev).eval()
rng = np.random.Generator(np.random.PCG64(1))
dense = torch.from_numpy((rng.random((B,) + obs_shape) < 0.5).astype(np.float32)).to(dev)
bits = pack_obs_bits(dense.cpu()).to(dev)
tcp = TensorCorePolicy(pol, max_batch=B, device=dev, with_value=True)
probs = torch.zeros((B, A), dtype=torch.float32, device=dev); logits = torch.zeros_like(probs); values = torch.zeros(B, dtype=torch.float32, device=dev)
tcp.forward_bits(bits, probs=probs, logits=logits, values=values)
torch.cuda.synchronize()
with torch.no_grad():
    ref_logits, ref_v = pol.double()(dense[:4096].double())
    pol.float()
err = float((logits[:4096].double() - ref_logits).abs().max())
rel = float(((logits[:4096].double() - ref_logits).abs() / ref_logits.abs().clamp_min(1e-3)).max())
out = {"batch": B, "max_abs_logit_err": err, "max_rel_logit_err": rel, "value_err": float((values[:4096].double() - ref_v.reshape(-1)).abs().max())}


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


out["tc_us"] = timed(lambda: tcp.forward_bits(bits, probs=probs, values=values))
out["tc_fused_kernel"] = tcp.set_per_layer(True) is False and tcp.set_per_layer(False)
tcp.set_per_layer(True)
out["tc_per_layer_us"] = timed(lambda: tcp.forward_bits(bits, probs=probs, values=values))
tcp.set_per_layer(False)


def torch_fwd(prec):
    def f():
        with torch.no_grad():
            if prec == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    l, v = pol(dense)
            else:
                torch.backends.cuda.matmul.allow_tf32 = prec == "tf32"
                l, v = pol(dense)
            torch.softmax(l.float(), dim=-1)
    return f


for prec in ("f32", "tf32", "bf16"):
    out[f"torch_{prec}_us"] = timed(torch_fwd(prec))
torch.backends.cuda.matmul.allow_tf32 = False
flops = 2 * B * (256 * 512 * 2 + 512 * 256 * 3 + 256 * 80 * 3)
out["tc_tflops_f16_products"] = flops / (out["tc_us"] * 1e-6) / 1e12
print(json.dumps(out))

# tools build with -DQG_TC_PROBE: where the MMA-issuing threads wait
try:
    import ctypes as C
    from qiskit_gym_b200._lib import lib
    L = lib()
    if hasattr(L, "qg_policy_tc_debug_read"):
        tcp.forward_bits(bits, probs=probs, values=values); torch.cuda.synchronize()
        buf = (C.c_longlong * (160 * 8))()
        L.qg_policy_tc_debug_read(buf)
        w = np.array(list(buf), dtype=np.float64).reshape(160, 8)[:148]
        names = ["d1_empty", "full_l1", "a0_full", "d2_empty", "act_full_l2", "full_l23", "act_full_l3", "total"]
        print(json.dumps({"mma_thread_wait_cycles_mean": {n: float(w[:, i].mean()) for i, n in enumerate(names)},
                          "share_of_total": {n: float(w[:, i].mean() / w[:, 7].mean()) for i, n in enumerate(names)}}))
        if hasattr(L, "qg_policy_tc_debug_read_epilogue"):
            L.qg_policy_tc_debug_read_epilogue(buf)
            w = np.array(list(buf), dtype=np.float64).reshape(160, 8)[:148]
            names = ["wait_d1_full", "wait_act_empty", "convert_l1", "wait_d2_full", "l2_pieces", "head_wait", "head", "total"]
            print(json.dumps({"epilogue_warp2_cycles_mean": {n: float(w[:, i].mean()) for i, n in enumerate(names)},
                              "share_of_total": {n: float(w[:, i].mean() / w[:, 7].mean()) for i, n in enumerate(names)}}))
except Exception as ex:  # noqa: BLE001
    print("probe read failed:", ex, file=sys.stderr)
