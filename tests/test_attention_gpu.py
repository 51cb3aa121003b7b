"""GPU parity: packed varlen causal attention (TMA-staged, mma.sync) vs the oracle.

Tolerance: inputs are bf16; P and dS are rounded to bf16 before the second matmul (as flash-attn
does), so outputs agree with the fp32 oracle on the same bf16 inputs to ~1e-2 absolute for unit-
variance data (the reference allows 1e-2 between eager and flash-attn on loss/grad-norm,
tests/models/test_models_patch.py:327-329; SP attention 2e-3 on grads of fp32 SDPA).
"""
import math

import os

import pytest
import torch

from oracle import attention as o_attn

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _run(cuda_dev, lens, Hq, Hk, D, seed=0, causal=True, bwd=True, atol=2e-2):
    from veomni_b200.attention import flash_attn_varlen

    g = torch.Generator().manual_seed(seed)
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    q = torch.randn(T, Hq, D, generator=g).to(BF)
    k = torch.randn(T, Hk, D, generator=g).to(BF)
    v = torch.randn(T, Hk, D, generator=g).to(BF)
    qg, kg, vg = (t.to(cuda_dev).requires_grad_(True) for t in (q, k, v))
    o = flash_attn_varlen(qg, kg, vg, cu.to(cuda_dev), max(lens), None, causal)
    o_ref, _ = o_attn.varlen_causal_attention(q.float(), k.float(), v.float(), cu, causal=causal)
    torch.testing.assert_close(o.float().cpu(), o_ref, atol=atol, rtol=2e-2)
    if bwd:
        do = torch.randn(T, Hq, D, generator=g).to(BF)
        o.backward(do.to(cuda_dev))
        dq, dk, dv = o_attn.varlen_causal_attention_bwd(q.float(), k.float(), v.float(), cu, do.float(), causal=causal)
        for name, got, ref in (("dq", qg.grad, dq), ("dk", kg.grad, dk), ("dv", vg.grad, dv)):
            scale = max(1.0, float(ref.abs().max()))
            torch.testing.assert_close(got.float().cpu() / scale, ref / scale, atol=atol, rtol=3e-2,
                                       msg=lambda m, name=name: f"{name}: {m}")
    return o


def test_attention_matches_reference_fixture(cuda_dev, golden):
    from veomni_b200.attention import flash_attn_varlen

    f = golden("ops.pt")["attention/fp32"]
    q, k, v = (f[n].to(BF) for n in ("q", "k", "v"))
    o = flash_attn_varlen(q.to(cuda_dev), k.to(cuda_dev), v.to(cuda_dev), f["cu"].to(cuda_dev), 14)
    # the fixture is the reference's eager attention on the fp32 inputs; bf16 input rounding dominates
    torch.testing.assert_close(o.float().cpu(), f["out"], atol=4e-2, rtol=4e-2)
    o_ref, _ = o_attn.varlen_causal_attention(q.float(), k.float(), v.float(), f["cu"])
    torch.testing.assert_close(o.float().cpu(), o_ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("lens", [[1], [7], [64], [65], [128], [129], [1, 63, 64, 65, 127, 129, 300], [257, 3, 511]])
@pytest.mark.parametrize("D", [64, 128])
def test_attention_ragged_lengths(cuda_dev, lens, D):
    _run(cuda_dev, lens, Hq=4, Hk=2, D=D, seed=len(lens) + D)


@pytest.mark.parametrize("Hq,Hk", [(1, 1), (8, 8), (8, 2), (6, 1)])
def test_attention_gqa(cuda_dev, Hq, Hk):
    _run(cuda_dev, [200, 333], Hq, Hk, 128, seed=Hq * 10 + Hk)


def test_attention_non_causal(cuda_dev):
    _run(cuda_dev, [100, 260], 4, 2, 128, seed=5, causal=False)


def test_attention_strided_qkv_views(cuda_dev):
    """q/k/v as column slices of a fused projection output (no copy on the way in)."""
    from veomni_b200.attention import flash_attn_varlen

    g = torch.Generator().manual_seed(11)
    T, Hq, Hk, D = 300, 4, 2, 128
    qkv = torch.randn(T, (Hq + 2 * Hk) * D, generator=g).to(BF)
    dev_qkv = qkv.to(cuda_dev)
    q = dev_qkv[:, : Hq * D].view(T, Hq, D)
    k = dev_qkv[:, Hq * D : (Hq + Hk) * D].view(T, Hk, D)
    v = dev_qkv[:, (Hq + Hk) * D :].view(T, Hk, D)
    cu = torch.tensor([0, 120, 300], dtype=torch.int32)
    o = flash_attn_varlen(q, k, v, cu.to(cuda_dev), 180)
    o_ref, _ = o_attn.varlen_causal_attention(q.float().cpu(), k.float().cpu(), v.float().cpu(), cu)
    torch.testing.assert_close(o.float().cpu(), o_ref, atol=2e-2, rtol=2e-2)


def test_attention_full_size_qwen3_8b_shape(cuda_dev):
    """BASELINE size (T=4096, 32/8 heads, D=128): one GQA group against the oracle + determinism + causality."""
    from veomni_b200.attention import flash_attn_varlen

    g = torch.Generator().manual_seed(42)
    T, Hq, Hk, D = 4096, 32, 8, 128
    q = torch.randn(T, Hq, D, generator=g).to(BF)
    k = torch.randn(T, Hk, D, generator=g).to(BF)
    v = torch.randn(T, Hk, D, generator=g).to(BF)
    do = torch.randn(T, Hq, D, generator=g).to(BF)
    cu = torch.tensor([0, T], dtype=torch.int32).to(cuda_dev)

    def run():
        qg, kg, vg = (t.to(cuda_dev).requires_grad_(True) for t in (q, k, v))
        o = flash_attn_varlen(qg, kg, vg, cu, T)
        o.backward(do.to(cuda_dev))
        return o.detach(), qg.grad, kg.grad, vg.grad

    o1, dq1, dk1, dv1 = run()
    o2, dq2, dk2, dv2 = run()
    for a, b in ((o1, o2), (dq1, dq2), (dk1, dk2), (dv1, dv2)):
        assert torch.equal(a, b), "attention forward/backward must be bit-reproducible (no atomics)"
    # oracle on kv head 3 and its 4 q heads
    hs = slice(12, 16)
    o_ref, _ = o_attn.varlen_causal_attention(q[:, hs].float(), k[:, 3:4].float(), v[:, 3:4].float(), cu.cpu())
    torch.testing.assert_close(o1[:, hs].float().cpu(), o_ref, atol=2e-2, rtol=2e-2)
    dq, dk, dv = o_attn.varlen_causal_attention_bwd(q[:, hs].float(), k[:, 3:4].float(), v[:, 3:4].float(), cu.cpu(),
                                                    do[:, hs].float())
    # dk/dv of kv head 3 only see its own q heads
    for name, got, ref in (("dq", dq1[:, hs], dq), ("dk", dk1[:, 3:4], dk), ("dv", dv1[:, 3:4], dv)):
        s = max(1.0, float(ref.abs().max()))
        torch.testing.assert_close(got.float().cpu() / s, ref / s, atol=2e-2, rtol=3e-2, msg=lambda m, n=name: f"{n}: {m}")
    # causality: the first 1000 outputs do not depend on later tokens
    cu2 = torch.tensor([0, 1000], dtype=torch.int32).to(cuda_dev)
    o_short = flash_attn_varlen(q[:1000].to(cuda_dev), k[:1000].to(cuda_dev), v[:1000].to(cuda_dev), cu2, 1000)
    torch.testing.assert_close(o_short.float(), o1[:1000].float(), atol=1e-2, rtol=1e-2)


@pytest.fixture
def tc_forward():
    from veomni_b200 import attention as A

    old = A.FWD_IMPL
    A.FWD_IMPL = "tc"
    yield
    A.FWD_IMPL = old


@pytest.mark.parametrize("lens", [[1], [64], [127], [128], [129], [1, 63, 64, 65, 127, 129, 300], [257, 3, 511], [1000, 24]])
def test_attention_tcgen05_forward_ragged(cuda_dev, tc_forward, lens):
    """tcgen05/TMEM forward (attention_tc.cu) against the oracle; backward runs on its saved o/lse."""
    _run(cuda_dev, lens, Hq=4, Hk=2, D=128, seed=len(lens) + 7)


def test_attention_tcgen05_forward_gqa_noncausal(cuda_dev, tc_forward):
    _run(cuda_dev, [200, 333], 8, 2, 128, seed=3)
    _run(cuda_dev, [100, 260], 4, 4, 128, seed=4, causal=False)


def test_attention_tcgen05_matches_mma_at_full_size(cuda_dev):
    from veomni_b200 import attention as A

    g = torch.Generator().manual_seed(9)
    T, Hq, Hk, D = 4096, 32, 8, 128
    q, k, v = (torch.randn(T, h, D, generator=g).to(BF).to(cuda_dev) for h in (Hq, Hk, Hk))
    cu = torch.tensor([0, 1500, T], dtype=torch.int32, device=cuda_dev)
    old = A.FWD_IMPL
    try:
        A.FWD_IMPL = "mma"
        o1, l1 = A.flash_attn_varlen(q, k, v, cu, 2596, return_lse=True)
        A.FWD_IMPL = "tc"
        o2, l2 = A.flash_attn_varlen(q, k, v, cu, 2596, return_lse=True)
        o3, _ = A.flash_attn_varlen(q, k, v, cu, 2596, return_lse=True)
    finally:
        A.FWD_IMPL = old
    assert torch.equal(o2, o3), "tcgen05 forward must be deterministic"
    torch.testing.assert_close(o2.float(), o1.float(), atol=1e-2, rtol=2e-2)
    torch.testing.assert_close(l2, l1, atol=2e-3, rtol=1e-4)


@pytest.fixture
def tc_backward():
    from veomni_b200 import attention as A

    old = (A.FWD_IMPL, A.BWD_IMPL)
    A.FWD_IMPL, A.BWD_IMPL = "tc", "tc"
    yield
    A.FWD_IMPL, A.BWD_IMPL = old


@pytest.mark.parametrize("lens", [[1], [64], [65], [127], [128], [129], [1, 63, 64, 65, 127, 129, 300], [257, 3, 511], [1000, 24]])
def test_attention_tcgen05_backward_ragged(cuda_dev, tc_backward, lens):
    _run(cuda_dev, lens, Hq=4, Hk=2, D=128, seed=len(lens) + 11)


def test_attention_tcgen05_backward_gqa_noncausal(cuda_dev, tc_backward):
    _run(cuda_dev, [200, 333], 8, 2, 128, seed=13)
    _run(cuda_dev, [200, 333], 6, 1, 128, seed=14)
    _run(cuda_dev, [100, 260], 4, 4, 128, seed=15, causal=False)


def test_attention_tcgen05_backward_matches_mma_full_size(cuda_dev):
    from veomni_b200 import attention as A

    g = torch.Generator().manual_seed(21)
    T, Hq, Hk, D = 4096, 32, 8, 128
    q, k, v, do = (torch.randn(T, h, D, generator=g).to(BF).to(cuda_dev) for h in (Hq, Hk, Hk, Hq))
    cu = torch.tensor([0, T], dtype=torch.int32, device=cuda_dev)
    res = {}
    old = (A.FWD_IMPL, A.BWD_IMPL)
    try:
        for impl in ("mma", "tc", "tc2"):
            A.FWD_IMPL = A.BWD_IMPL = "tc" if impl.startswith("tc") else "mma"
            qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
            A.flash_attn_varlen(qq, kk, vv, cu, T).backward(do)
            res[impl] = (qq.grad, kk.grad, vv.grad)
    finally:
        A.FWD_IMPL, A.BWD_IMPL = old
    for a, b in zip(res["tc"], res["tc2"]):
        assert torch.equal(a, b), "tcgen05 backward must be deterministic"
    for name, a, b in zip(("dq", "dk", "dv"), res["tc"], res["mma"]):
        s = max(1.0, float(b.float().abs().max()))
        torch.testing.assert_close(a.float() / s, b.float() / s, atol=1e-2, rtol=2e-2, msg=lambda m, n=name: f"{n}: {m}")


def test_attention_tcgen05_operand_variants_agree(cuda_dev):
    """The backward kernels with their A operands in TMEM (default: Q / dO resident for dQ; P^T / dS^T for dK, dV) and the
    all-shared-memory-operand variants issue the same MMAs on the same data: bit-identical gradients, on ragged lengths
    and at the Qwen3-8B size."""
    from veomni_b200 import attention as A

    old = (A.FWD_IMPL, A.BWD_IMPL, A.BWD_DQ_SS, A.BWD_PP, A.BWD_P16, A.BWD_DQ_N128)
    try:
        A.FWD_IMPL = A.BWD_IMPL = "tc"
        for lens, Hq, Hk in (([1, 63, 64, 65, 127, 129, 300], 4, 2), ([700, 64], 6, 1), ([4096], 32, 8)):
            T = sum(lens)
            g = torch.Generator().manual_seed(T)
            q, k, v, do = (torch.randn(T, h, 128, generator=g).to(BF).to(cuda_dev) for h in (Hq, Hk, Hk, Hq))
            cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
            grads = {}
            for variant in ("tmem", "smem", "pingpong", "p16", "dq_n128"):
                A.BWD_DQ_SS, A.BWD_PP, A.BWD_P16 = variant == "smem", variant == "pingpong", variant == "p16"
                A.BWD_DQ_N128 = variant == "dq_n128"
                qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
                A.flash_attn_varlen(qq, kk, vv, cu, max(lens)).backward(do)
                grads[variant] = (qq.grad, kk.grad, vv.grad)
            for variant in ("smem", "pingpong", "p16", "dq_n128"):
                for a, b in zip(grads["tmem"], grads[variant]):
                    assert torch.equal(a, b), (variant, lens)
    finally:
        A.FWD_IMPL, A.BWD_IMPL, A.BWD_DQ_SS, A.BWD_PP, A.BWD_P16, A.BWD_DQ_N128 = old


def test_attention_tcgen05_forward_eight_softmax_warps(cuda_dev):
    """attn_fwd_tc_kernel<W8> (two warps per row, half the columns each) against the default four-warp kernel: the row
    maxima, the lazy-rescale decisions and every exponential are the same numbers, only the row sums are added in a
    different order -> outputs within 1 bf16 ulp, lse within 1e-5; plus the oracle parity of the ragged cases."""
    from veomni_b200 import attention as A

    old = (A.FWD_IMPL, A.FWD_W8)
    try:
        A.FWD_IMPL = "tc"
        for lens, Hq, Hk, causal in (([1, 63, 64, 65, 127, 129, 300], 4, 2, True), ([200, 333], 8, 2, False), ([4096], 32, 8, True)):
            T = sum(lens)
            g = torch.Generator().manual_seed(T + 1)
            q, k, v = (torch.randn(T, h, 128, generator=g).to(BF).to(cuda_dev) for h in (Hq, Hk, Hk))
            cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
            outs = {}
            for w8 in (False, True):
                A.FWD_W8 = w8
                with torch.no_grad():
                    outs[w8] = A.flash_attn_varlen(q, k, v, cu, max(lens), causal=causal, return_lse=True)
            torch.testing.assert_close(outs[True][0].float(), outs[False][0].float(), atol=1e-2, rtol=8e-3)
            torch.testing.assert_close(outs[True][1], outs[False][1], atol=1e-5, rtol=1e-5)
        A.FWD_W8 = True
        _run(cuda_dev, [1, 63, 64, 65, 127, 129, 300], Hq=4, Hk=2, D=128, seed=99)
    finally:
        A.FWD_IMPL, A.FWD_W8 = old
