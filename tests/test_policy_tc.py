"""The tensor-core policy (qg_policy_tc_*, csrc/qg_policy_tc.cu: tcgen05.mma, f16 hi+lo split operands, f32 accumulation in tensor memory)
against a float64 evaluation of the PyTorch BasicPolicy it was built from — the same tolerance as the f32 FFMA kernel's test
(tests/test_policy.py): logits within 1e-4 relative / 2e-5 absolute, softmax within 1e-4 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("obs_shape,A,emb,common,density", [
    ((16, 16), 72, 512, (256,), 0.5),          # C3 CliffordGym 8q: the collector's network
    ((27, 27), 28, 512, (256,), 0.04),         # C5 PermutationGym 27q (one-hot rows)
    ((9, 9), 12, 64, (32,), 0.12),             # widths that are not multiples of the K block
    ((20, 25), 104, 300, (200, 100), 0.3),     # three hidden layers, 64-wide column tiles
    ((3, 3), 2, 1024, (), 0.5),
])
def test_tensor_core_policy_matches_float64(obs_shape, A, emb, common, density):
    from qiskit_gym_b200.policy import TensorCorePolicy, pack_obs_bits
    from qiskit_gym_b200.search import BasicPolicy
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    pol = BasicPolicy(list(obs_shape), A, embedding_size=emb, common_layers=common).to(dev).eval()
    tcp = TensorCorePolicy(pol, max_batch=5000, device=dev, with_value=True)
    rng = np.random.Generator(np.random.PCG64(9))
    for B in (1, 127, 128, 129, 5000):
        dense = torch.from_numpy((rng.random((B,) + tuple(obs_shape)) < density).astype(np.float32))
        if B > 2:
            dense[0] = 0.0
            dense[1] = 1.0
        bits = pack_obs_bits(dense).to(dev)
        probs = torch.full((B, A), -1.0, dtype=torch.float32, device=dev)
        logits = torch.full((B, A), -1.0, dtype=torch.float32, device=dev)
        values = torch.full((B,), -1.0, dtype=torch.float32, device=dev)
        tcp.forward_bits(bits, probs=probs, logits=logits, values=values)
        with torch.no_grad():
            ref_logits, ref_v = pol.double()(dense.to(dev).double())
            pol.float()
        ref_probs = torch.softmax(ref_logits, dim=-1)
        assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=2e-5), (B, float((logits.double() - ref_logits).abs().max()))
        assert torch.allclose(values.double(), ref_v.reshape(-1), rtol=1e-4, atol=2e-5), (B, float((values.double() - ref_v.reshape(-1)).abs().max()))
        assert torch.allclose(probs.double(), ref_probs, rtol=1e-4, atol=1e-6)
        assert torch.allclose(probs.sum(dim=1), torch.ones(B, device=dev), atol=1e-5)
    # the fused three-layer kernel (where the shape allows it) and the one-kernel-per-layer path do the same products in the same order
    fused_runs = tcp.set_per_layer(False)
    assert fused_runs == (len(common) == 1 and ((emb + 63) // 64 * 64) % 128 == 0 and emb <= 512 and (common[0] + 63) // 64 * 64 <= 256)
    if fused_runs:
        tcp.set_per_layer(True)
        l2 = torch.empty_like(logits); p2 = torch.empty_like(probs); v2 = torch.empty_like(values)
        tcp.forward_bits(bits, probs=p2, logits=l2, values=v2)
        assert torch.allclose(l2, logits, rtol=0, atol=1e-6) and torch.allclose(p2, probs, rtol=0, atol=1e-7) and torch.allclose(v2, values, rtol=0, atol=1e-6)
        assert torch.allclose(l2.double(), ref_logits, rtol=1e-4, atol=2e-5)
        tcp.set_per_layer(False)
    # a second call on the same handle (buffers are reused) and a handle without the value head
    tcp.forward_bits(bits, logits=logits)
    assert torch.allclose(logits.double(), ref_logits, rtol=1e-4, atol=2e-5)
    tcp2 = TensorCorePolicy(pol, max_batch=256, device=dev, with_value=False)
    small = bits[:200].contiguous()
    out = tcp2.forward_bits(small)
    assert torch.allclose(out.double(), ref_probs[:200], rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError):
        tcp2.forward_bits(bits)                      # larger than max_batch
