"""FusedPolicy — the action network of a twisterl `BasicPolicy` evaluated from packed-bit observations by one CUDA kernel
(`qg_policy_forward_bits`, csrc/qg_policy.cu).

twisterl's policy consumes `Env::observe()`'s sparse indices with a gather-sum first layer; the PyTorch `BasicPolicy` in
search.py consumes the dense f32 tensor instead.  `FusedPolicy` takes the weights of such a module (or of a checkpoint in
the reference's `.pt` format) and evaluates  obs bits -> Linear -> ReLU -> ... -> Linear -> softmax  in a single launch,
which is what makes a synth-search decision two launches long.  Its output matches the PyTorch module to f32 rounding
(sums run in a different order than cuBLAS'); tests/test_policy.py states the tolerance.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from .engine import _dptr


def action_layers(policy: torch.nn.Module):
    """The Linear layers on the observation -> action-logits path of a BasicPolicy-shaped module, in order."""
    layers = [policy.embeddings]
    layers += [m for m in policy.common if isinstance(m, torch.nn.Linear)]
    layers += [m for m in policy.action if isinstance(m, torch.nn.Linear)]
    return layers


def simple_value_head(policy: torch.nn.Module):
    """The value head as (weight [in], bias) when it is a single Linear fed by the same activations as a single-Linear action head (the
    reference's default BasicPolicy and all of its example checkpoints: policy_layers = value_layers = []), else None."""
    act = [m for m in policy.action if isinstance(m, torch.nn.Linear)]
    val = [m for m in getattr(policy, "value", []) if isinstance(m, torch.nn.Linear)]
    if len(act) == 1 and len(val) == 1 and val[0].out_features == 1 and val[0].in_features == act[0].in_features:
        return val[0]
    return None


def weights_version(policy: torch.nn.Module) -> int:
    """Changes whenever a parameter of the module is updated in place (optimizer steps, load_state_dict)."""
    return sum(int(p._version) for p in policy.parameters())


class FusedPolicy:
    def __init__(self, policy: torch.nn.Module, device=None, with_value: bool = False):
        """with_value: also evaluate the value head (forward_bits(..., values=...)); needs the simple head layout (simple_value_head)."""
        if not torch.cuda.is_available():
            raise RuntimeError("qiskit_gym_b200 needs a CUDA device (no CPU fallback)")
        self.device_index = torch.cuda.current_device() if device is None else (device.index if isinstance(device, torch.device) else int(device))
        self.device = torch.device("cuda", self.device_index)
        layers = action_layers(policy)
        ws = [np.ascontiguousarray(l.weight.detach().cpu().numpy(), dtype=np.float32) for l in layers]
        bs = [np.ascontiguousarray(l.bias.detach().cpu().numpy(), dtype=np.float32) for l in layers]
        for a, b in zip(ws[:-1], ws[1:]):
            assert b.shape[1] == a.shape[0], "layers do not chain"
        self.obs_size = int(ws[0].shape[1])
        self.obs_words = (self.obs_size + 31) // 32
        self.num_actions = int(ws[-1].shape[0])
        self.version = weights_version(policy)
        n = len(ws)
        widths = (C.c_int32 * n)(*[int(w.shape[0]) for w in ws])
        wp = (C.c_void_p * n)(*[w.ctypes.data for w in ws])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        h = C.c_void_p()
        self.has_value = False
        if with_value:
            head = simple_value_head(policy)
            if head is None:
                raise NotImplementedError("FusedPolicy(with_value=True) needs single-Linear action and value heads on the same activations")
            vw = np.ascontiguousarray(head.weight.detach().cpu().numpy().reshape(-1), dtype=np.float32)
            vb = float(head.bias.detach().cpu().numpy().reshape(-1)[0])
            check(lib().qg_policy_create_value(self.device_index, self.obs_size, n, widths, wp, bp, vw.ctypes.data_as(C.c_void_p), C.c_float(vb), C.byref(h)))
            self.has_value = True
        else:
            check(lib().qg_policy_create(self.device_index, self.obs_size, n, widths, wp, bp, C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().qg_policy_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward_bits(self, obs_bits: torch.Tensor, probs: torch.Tensor | None = None, logits: torch.Tensor | None = None,
                     values: torch.Tensor | None = None):
        """obs_bits int32 [B, obs_words] (BatchedEnv.observe_bits / step_bits) -> softmax action weights f32 [B, A]
        (and the value head's output f32 [B] into `values` when given)."""
        assert obs_bits.is_cuda and obs_bits.element_size() == 4 and obs_bits.is_contiguous() and obs_bits.shape[-1] == self.obs_words
        B = obs_bits.numel() // self.obs_words
        if probs is None and logits is None and values is None:
            probs = torch.empty((B, self.num_actions), dtype=torch.float32, device=obs_bits.device)
        for t in (probs, logits):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == B * self.num_actions)
        assert values is None or (self.has_value and values.dtype == torch.float32 and values.is_contiguous() and values.numel() == B)
        st = C.c_void_p(torch.cuda.current_stream(obs_bits.device).cuda_stream)
        if values is None:
            check(lib().qg_policy_forward_bits(self._h, _dptr(obs_bits), B, _dptr(probs), _dptr(logits), st))
        else:
            check(lib().qg_policy_forward_bits_value(self._h, _dptr(obs_bits), B, _dptr(probs), _dptr(logits), _dptr(values), st))
        return probs if probs is not None else (logits if logits is not None else values)


class TensorCorePolicy:
    """The same network for large batches on tcgen05 tensor cores (`qg_policy_tc_*`, csrc/qg_policy_tc.cu): packed observation bits in,
    softmax action weights / logits / values out; f16 hi+lo split operands with f32 accumulation (logits within 1e-4 of the f32 module)."""

    def __init__(self, policy: torch.nn.Module, max_batch: int, device=None, with_value: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("qiskit_gym_b200 needs a CUDA device (no CPU fallback)")
        self.device_index = torch.cuda.current_device() if device is None else (device.index if isinstance(device, torch.device) else int(device))
        self.device = torch.device("cuda", self.device_index)
        layers = action_layers(policy)
        ws = [np.ascontiguousarray(l.weight.detach().cpu().numpy(), dtype=np.float32) for l in layers]
        bs = [np.ascontiguousarray(l.bias.detach().cpu().numpy(), dtype=np.float32) for l in layers]
        self.obs_size = int(ws[0].shape[1])
        self.obs_words = (self.obs_size + 31) // 32
        self.num_actions = int(ws[-1].shape[0])
        self.max_batch = int(max_batch)
        self.version = weights_version(policy)
        n = len(ws)
        widths = (C.c_int32 * n)(*[int(w.shape[0]) for w in ws])
        wp = (C.c_void_p * n)(*[w.ctypes.data for w in ws])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        vw, vb = None, 0.0
        self.has_value = False
        if with_value:
            head = simple_value_head(policy)
            if head is None:
                raise NotImplementedError("TensorCorePolicy(with_value=True) needs single-Linear action and value heads on the same activations")
            vwa = np.ascontiguousarray(head.weight.detach().cpu().numpy().reshape(-1), dtype=np.float32)
            vw, vb = vwa.ctypes.data_as(C.c_void_p), float(head.bias.detach().cpu().numpy().reshape(-1)[0])
            self.has_value = True
        h = C.c_void_p()
        check(lib().qg_policy_tc_create(self.device_index, self.obs_size, n, widths, wp, bp, vw, C.c_float(vb), self.max_batch, C.byref(h)))
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().qg_policy_tc_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_per_layer(self, per_layer: bool) -> bool:
        """Forces the one-kernel-per-layer path (True) or lets the fused three-layer kernel run where it applies (False); returns whether
        the fused kernel will run."""
        return bool(lib().qg_policy_tc_set_mode(self._h, 1 if per_layer else 0))

    def forward_bits(self, obs_bits: torch.Tensor, probs: torch.Tensor | None = None, logits: torch.Tensor | None = None,
                     values: torch.Tensor | None = None):
        assert obs_bits.is_cuda and obs_bits.element_size() == 4 and obs_bits.is_contiguous() and obs_bits.shape[-1] == self.obs_words
        B = obs_bits.numel() // self.obs_words
        if probs is None and logits is None and values is None:
            probs = torch.empty((B, self.num_actions), dtype=torch.float32, device=obs_bits.device)
        for t in (probs, logits):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == B * self.num_actions)
        assert values is None or (self.has_value and values.dtype == torch.float32 and values.is_contiguous() and values.numel() == B)
        st = C.c_void_p(torch.cuda.current_stream(obs_bits.device).cuda_stream)
        check(lib().qg_policy_tc_forward_bits(self._h, _dptr(obs_bits), B, _dptr(probs), _dptr(logits), _dptr(values), st))
        return probs if probs is not None else (logits if logits is not None else values)


def pack_obs_bits(obs: torch.Tensor) -> torch.Tensor:
    """Dense 0/1 observation [B, ...] -> packed int32 [B, ceil(obs/32)] (host-side helper for tests and tools)."""
    flat = (obs.reshape(obs.shape[0], -1) != 0).cpu().numpy()
    B, n = flat.shape
    words = (n + 31) // 32
    pad = np.zeros((B, words * 32), dtype=bool)
    pad[:, :n] = flat
    packed = np.packbits(pad.reshape(B, words, 32), axis=2, bitorder="little").view(np.uint32).reshape(B, words)
    return torch.from_numpy(packed.view(np.int32).copy())
