"""GPU parity: MoE routing (bit-exact), row scatter/gather, tcgen05 GroupGEMM, fused MoE fwd/bwd vs the oracle.

Tolerances follow the reference's own fused-vs-eager bars (tests/ops/test_fused_moe_split_vs_merged.py:157-161):
forward 1e-2, hidden-state grad 5e-2, weight grads 3e-2 on 0.1-scaled data.
"""
import pytest
import torch

from oracle import moe as o_moe

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rand_idx(T, K, E, seed, skew=False):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, E, generator=g)
    if skew:
        logits[:, : max(1, E // 8)] += 3.0
    return torch.topk(logits, K, dim=-1).indices


@pytest.mark.parametrize("T,K,E,dtype", [(1, 1, 4, torch.int64), (48, 2, 8, torch.int64), (1000, 8, 128, torch.int64),
                                         (4096, 8, 128, torch.int64), (8192, 8, 128, torch.int32), (300, 6, 64, torch.int32),
                                         (1024, 1, 1024, torch.int64)])
def test_route_indices_bit_exact(cuda_dev, T, K, E, dtype):
    from veomni_b200.moe import moe_route

    idx = _rand_idx(T, K, E, T + K, skew=True).to(dtype)
    splits, cumsum, sidx = moe_route(idx.to(cuda_dev), E)
    assert torch.equal(splits.cpu(), o_moe.expert_histogram(idx, E))
    assert torch.equal(cumsum.cpu().long(), torch.cumsum(o_moe.expert_histogram(idx, E).long(), 0))
    assert torch.equal(sidx.cpu(), o_moe.scatter_index(idx))  # == argsort(stable).argsort(), group_gemm.py:44


def test_route_edge_cases(cuda_dev, golden):
    from veomni_b200.moe import moe_route

    # all slots on one expert; empty experts (reference test_quack_fused_moe.py test_all_same_expert)
    idx = torch.full((513, 4), 3, dtype=torch.int64)
    splits, cumsum, sidx = moe_route(idx.to(cuda_dev), 16)
    assert torch.equal(sidx.cpu().flatten(), torch.arange(513 * 4, dtype=torch.int32))
    assert int(splits[3]) == 513 * 4 and int(splits.sum()) == 513 * 4
    # fixture produced next to the reference run (expert 5 empty)
    f = golden("moe.pt")["moe/fp32"]
    splits, cumsum, sidx = moe_route(f["idx"].to(cuda_dev), f["gate_up"].shape[0])
    assert torch.equal(sidx.cpu(), f["scatter_index"]) and torch.equal(splits.cpu(), f["splits"])
    # empty input
    splits, cumsum, sidx = moe_route(torch.empty(0, 2, dtype=torch.int64, device=cuda_dev), 8)
    assert int(splits.sum()) == 0 and sidx.numel() == 0


@pytest.mark.parametrize("T,K,H", [(5, 2, 64), (257, 8, 2048), (4096, 8, 2048)])
def test_scatter_gather_vs_oracle(cuda_dev, T, K, H):
    from veomni_b200.moe import moe_gather, moe_route, moe_scatter

    g = torch.Generator().manual_seed(T)
    idx = _rand_idx(T, K, 32, T)
    x = torch.randn(T, H, generator=g).to(BF)
    _, _, sidx = moe_route(idx.to(cuda_dev), 32)
    xg = x.to(cuda_dev).requires_grad_(True)
    s = moe_scatter(xg, sidx)
    assert torch.equal(s.cpu(), o_moe.moe_scatter(x, sidx.cpu()))  # pure copy: bit-exact
    y = torch.randn(T * K, H, generator=g).to(BF)
    out = moe_gather(y.to(cuda_dev), sidx)
    assert torch.equal(out.cpu(), o_moe.moe_gather(y, sidx.cpu()))  # fp32 acc in k order, one rounding: bit-exact
    # autograd pairing: d(scatter) = gather
    s.backward(y.to(cuda_dev))
    assert torch.equal(xg.grad.cpu(), o_moe.moe_gather(y, sidx.cpu()))


@pytest.mark.parametrize("T,K,H", [(3, 2, 256), (257, 8, 2048), (1000, 6, 4096), (64, 2, 64)])
def test_routing_weight_grad_vs_fp32_einsum(cuda_dev, T, K, H):
    """d(loss)/d(routing weight) of the EP combine: <g[t], rows[sidx[t,k]]> in fp32 (H = 64 takes the torch expression)."""
    from veomni_b200.ep import routing_weight_grad
    from veomni_b200.moe import moe_route

    gen = torch.Generator().manual_seed(T + H)
    idx = _rand_idx(T, K, 32, T + 1)
    _, _, sidx = moe_route(idx.to(cuda_dev), 32)
    g = torch.randn(T, H, generator=gen).to(BF)
    rows = torch.randn(T * K, H, generator=gen).to(BF)
    got = routing_weight_grad(g.to(cuda_dev), rows.to(cuda_dev), sidx).cpu()
    want = torch.einsum("th,tkh->tk", g.double(), rows[sidx.cpu().flatten().long()].view(T, K, H).double()).float()
    assert got.dtype == torch.float32 and got.shape == (T, K)
    torch.testing.assert_close(got, want, atol=1e-3 * H ** 0.5, rtol=1e-4)  # fp32 accumulation order only


def _ragged(G, total, seed, empties=()):
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(G, generator=g)
    for e in empties:
        w[e] = 0
    sizes = torch.floor(w / w.sum() * total).long()
    sizes[int(torch.argmax(w))] += total - int(sizes.sum())
    return torch.cumsum(sizes, 0).to(torch.int32)


@pytest.mark.parametrize("G,total,N,K", [(1, 128, 128, 64), (3, 200, 128, 64), (4, 37, 64, 128), (8, 1000, 256, 192),
                                         (16, 4096, 1536, 2048), (128, 4096, 2048, 768), (5, 300, 72, 40)])
@pytest.mark.parametrize("variant", [1, 2, 3])  # 1: tokens on the MMA M side, 2: tokens on N (default), 3: 2-CTA (cta_group::2) tokens on N
def test_group_gemm_nt_nn_vs_oracle(cuda_dev, G, total, N, K, variant, monkeypatch):
    from veomni_b200 import moe as M
    from veomni_b200.moe import group_gemm_same_nk

    monkeypatch.setattr(M, "GG_VARIANT", variant)
    cs = _ragged(G, total, G + total, empties=(1,) if G > 2 else ())
    g = torch.Generator().manual_seed(total)
    a = (0.5 * torch.randn(total + 7, K, generator=g)).to(BF)  # rows past cumsum[-1] are never written
    b_nt = (0.1 * torch.randn(G, N, K, generator=g)).to(BF)
    c = group_gemm_same_nk(a.to(cuda_dev), b_nt.to(cuda_dev), cs.to(cuda_dev), transpose_b=True)
    ref = o_moe.group_gemm_same_nk(a, b_nt, cs, transpose_b=True)
    torch.testing.assert_close(c[:total].float().cpu(), ref[:total].float(), atol=2e-2, rtol=2e-2)
    b_nn = b_nt.transpose(1, 2).contiguous()  # [G, K, N]
    c2 = group_gemm_same_nk(a.to(cuda_dev), b_nn.to(cuda_dev), cs.to(cuda_dev), transpose_b=False)
    torch.testing.assert_close(c2[:total].float().cpu(), ref[:total].float(), atol=2e-2, rtol=2e-2)
    assert torch.equal(c[:total], c2[:total])  # same products, same fp32 accumulation order


@pytest.mark.parametrize("G,total,M,N", [(1, 64, 128, 128), (3, 200, 128, 64), (4, 37, 64, 136), (8, 1000, 256, 192),
                                         (16, 2048, 1536, 2048), (128, 4096, 768, 2048)])
def test_group_gemm_tn_vs_oracle(cuda_dev, G, total, M, N):
    from veomni_b200.moe import group_gemm_same_mn

    cs = _ragged(G, total, G * 3 + total, empties=(0,) if G > 2 else ())
    g = torch.Generator().manual_seed(total + 1)
    a = (0.3 * torch.randn(total + 5, M, generator=g)).to(BF)
    b = (0.3 * torch.randn(total + 5, N, generator=g)).to(BF)
    c = torch.full((G, M, N), float("nan"), dtype=BF, device=cuda_dev)
    group_gemm_same_mn(a.to(cuda_dev), b.to(cuda_dev), c, cs.to(cuda_dev))
    ref = o_moe.group_gemm_same_mn(a, b, cs)
    s = max(1.0, float(ref.float().abs().max()))
    torch.testing.assert_close(c.float().cpu() / s, ref.float() / s, atol=2e-2, rtol=2e-2)
    if G > 2:
        assert float(c[0].float().abs().max()) == 0.0  # empty group is zero-filled (group_gemm.py:323-337)


def test_fused_moe_matches_reference_fixture(cuda_dev, golden):
    from veomni_b200.moe import fused_moe_forward

    f = golden("moe.pt")["moe/fp32"]
    E = f["gate_up"].shape[0]
    out = fused_moe_forward(E, f["rw"].to(BF).to(cuda_dev), f["idx"].to(cuda_dev), f["hs"].to(BF).to(cuda_dev), None, None,
                            f["down"].to(BF).to(cuda_dev), fc1_1_2_weight=f["gate_up"].to(BF).to(cuda_dev))
    torch.testing.assert_close(out.float().cpu(), f["y"], atol=1e-2, rtol=5e-2)  # reference eager fp32 output


@pytest.mark.parametrize("T,E,H,I,K", [(512, 128, 2048, 768, 8), (256, 64, 2048, 1408, 6), (64, 8, 128, 64, 2)])
def test_fused_moe_fwd_bwd_vs_eager(cuda_dev, T, E, H, I, K):
    """The reference's own test shapes (test_fused_moe_split_vs_merged.py:47-55): Qwen3-30B-A3B and Moonlight."""
    from veomni_b200.moe import fused_moe_forward

    g = torch.Generator().manual_seed(E + T)
    hs = (0.1 * torch.randn(T, H, generator=g)).to(BF)
    w1 = (0.1 * torch.randn(E, 2 * I, H, generator=g)).to(BF)
    w2 = (0.1 * torch.randn(E, H, I, generator=g)).to(BF)
    logits = torch.randn(T, E, generator=g)
    rw, idx = torch.topk(torch.softmax(logits, -1), K, dim=-1)
    rw = (rw / rw.sum(-1, keepdim=True)).to(BF)
    dy = (0.1 * torch.randn(T, H, generator=g)).to(BF)
    dev = [t.to(cuda_dev).requires_grad_(True) for t in (rw, hs, w1, w2)]
    out = fused_moe_forward(E, dev[0], idx.to(cuda_dev), dev[1], None, None, dev[3], fc1_1_2_weight=dev[2])
    out.backward(dy.to(cuda_dev))
    ref_in = [t.float().requires_grad_(True) for t in (rw, hs, w1, w2)]
    ref = o_moe.eager_moe_forward(E, ref_in[0], idx, ref_in[1], ref_in[2], ref_in[3])
    ref.backward(dy.float())
    torch.testing.assert_close(out.float().cpu(), ref.detach(), atol=1e-2, rtol=1e-2)
    for name, got, want, tol in (("d_routing", dev[0].grad, ref_in[0].grad, 5e-2), ("d_hidden", dev[1].grad, ref_in[1].grad, 5e-2),
                                 ("d_fc1", dev[2].grad, ref_in[2].grad, 3e-2), ("d_fc2", dev[3].grad, ref_in[3].grad, 3e-2)):
        s = max(1e-6, float(want.abs().max()))
        torch.testing.assert_close(got.float().cpu() / s, want / s, atol=tol, rtol=tol, msg=lambda m, n=name: f"{n}: {m}")
