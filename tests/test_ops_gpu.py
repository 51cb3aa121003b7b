"""GPU parity: memory-bound block ops through the C ABI vs the oracle / reference fixtures.

Tolerances (bf16 path): forward outputs within 1 bf16 ulp of the oracle evaluated on the same
bf16 inputs (atol/rtol 2^-7 relative, i.e. rtol=8e-3, plus a small atol) — the reference's own
registry test allows 2e-3..1e-2 between liger and eager (tests/ops/test_kernel_registry_numerical.py:70-136).
Gradients are compared with fp32 autograd of the oracle at rtol 2e-2.
"""
import pytest
import torch

from oracle import ops as o_ops

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _close(a, b, atol, rtol, what=""):
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), atol=atol, rtol=rtol, msg=lambda m: f"{what}: {m}")


def test_rmsnorm_fwd_matches_reference_fixture(cuda_dev, golden):
    from veomni_b200 import functional as F

    g = golden("ops.pt")
    for name in ("hidden", "head"):
        f = g[f"rms_norm/{name}/bf16"]
        y = F.rms_norm(f["x"].to(cuda_dev), f["w"].to(cuda_dev), f["eps"])
        # same rounding points as the reference => bit-exact up to rsqrt ulp flips
        _close(y, f["y"], atol=1e-2, rtol=8e-3, what=name)
        frac_exact = (y.cpu() == f["y"]).float().mean().item()
        assert frac_exact > 0.99, f"{name}: only {frac_exact:.4f} bit-exact"


@pytest.mark.parametrize("rows,cols", [(1, 8), (3, 64), (257, 128), (33, 264), (64, 896), (130, 2048), (4096, 4096), (17, 8192), (5, 16384)])
def test_rmsnorm_fwd_bwd_vs_oracle(cuda_dev, rows, cols):
    from veomni_b200 import functional as F

    g = torch.Generator().manual_seed(rows * 131 + cols)
    x = (torch.randn(rows, cols, generator=g) * 2).to(BF)
    w = (1 + 0.2 * torch.randn(cols, generator=g)).to(BF)
    dy = torch.randn(rows, cols, generator=g).to(BF)
    xg = x.to(cuda_dev).requires_grad_(True)
    wg = w.to(cuda_dev).requires_grad_(True)
    y = F.rms_norm(xg, wg, 1e-6)
    y.backward(dy.to(cuda_dev))
    y_ref = o_ops.rms_norm(x, w, 1e-6)
    _close(y, y_ref, atol=1e-2, rtol=8e-3, what="y")
    # fp32 autograd of the same function on the bf16-rounded inputs
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    o_ops.rms_norm(xf, wf, 1e-6).backward(dy.float())
    _close(xg.grad, xf.grad, atol=3e-2, rtol=2e-2, what="dx")
    scale = max(1.0, float(wf.grad.abs().max()))
    _close(wg.grad.float() / scale, wf.grad / scale, atol=2e-2, rtol=2e-2, what="dw")


def test_rmsnorm_empty_and_bad_shape(cuda_dev):
    from veomni_b200 import _lib
    from veomni_b200 import functional as F

    y = F.rms_norm(torch.empty(0, 128, dtype=BF, device=cuda_dev), torch.ones(128, dtype=BF, device=cuda_dev), 1e-6)
    assert y.shape == (0, 128)
    with pytest.raises(_lib.VB200Error):
        F.rms_norm(torch.ones(4, 100, dtype=BF, device=cuda_dev), torch.ones(100, dtype=BF, device=cuda_dev), 1e-6)
    with pytest.raises(_lib.VB200Error):
        F.rms_norm(torch.ones(4, 128, dtype=torch.float32, device=cuda_dev), torch.ones(128, device=cuda_dev), 1e-6)


def test_rope_matches_reference_fixture(cuda_dev, golden):
    from veomni_b200 import functional as F

    f = golden("ops.pt")["rope/bf16"]
    q, k = F.apply_rotary_pos_emb(f["q"].to(cuda_dev), f["k"].to(cuda_dev), f["cos"].to(cuda_dev), f["sin"].to(cuda_dev))
    assert q.shape == f["q_out"].shape and k.shape == f["k_out"].shape
    # reference rounds each product to bf16 before the add; we round once: <= 2 ulp apart
    _close(q, f["q_out"], atol=2e-2, rtol=1.6e-2, what="q")
    _close(k, f["k_out"], atol=2e-2, rtol=1.6e-2, what="k")


@pytest.mark.parametrize("S,Hq,Hk,D", [(5, 1, 1, 64), (77, 4, 2, 128), (130, 32, 8, 128), (9, 3, 1, 256), (4096, 32, 8, 128)])
def test_rope_fwd_bwd_vs_oracle(cuda_dev, S, Hq, Hk, D):
    from veomni_b200 import functional as F

    g = torch.Generator().manual_seed(S + D)
    # the reference layout: [B, S, H, D] projections viewed as [B, H, S, D]
    q = torch.randn(1, S, Hq, D, generator=g).to(BF).transpose(1, 2)
    k = torch.randn(1, S, Hk, D, generator=g).to(BF).transpose(1, 2)
    pos = torch.arange(S)[None]
    cos, sin = o_ops.rotary_cos_sin(pos, D, 1e6, BF)
    qg = q.to(cuda_dev).requires_grad_(True)
    kg = k.to(cuda_dev).requires_grad_(True)
    qo, ko = F.apply_rotary_pos_emb(qg, kg, cos.to(cuda_dev), sin.to(cuda_dev))
    qr, kr = o_ops.apply_rotary_pos_emb(q.float(), k.float(), cos.float(), sin.float())
    _close(qo, qr, atol=1e-2, rtol=8e-3, what="q")
    _close(ko, kr, atol=1e-2, rtol=8e-3, what="k")
    dq = torch.randn(qo.shape, generator=g).to(BF)
    dk = torch.randn(ko.shape, generator=g).to(BF)
    torch.autograd.backward([qo, ko], [dq.to(cuda_dev), dk.to(cuda_dev)])
    qf = q.float().requires_grad_(True)
    kf = k.float().requires_grad_(True)
    a, b = o_ops.apply_rotary_pos_emb(qf, kf, cos.float(), sin.float())
    torch.autograd.backward([a, b], [dq.float(), dk.float()])
    _close(qg.grad, qf.grad, atol=1e-2, rtol=8e-3, what="dq")
    _close(kg.grad, kf.grad, atol=1e-2, rtol=8e-3, what="dk")
    # rotation is orthogonal: norms are preserved (size-independent property)
    _close(qo.float().norm(dim=-1), q.float().norm(dim=-1), atol=5e-2, rtol=1e-2, what="norm")


@pytest.mark.parametrize("T,Hq,Hk,D", [(3, 2, 1, 64), (100, 4, 2, 128), (257, 32, 8, 128), (4096, 32, 8, 128), (50, 5, 3, 128)])
def test_qknorm_rope_fwd_bwd_vs_oracle(cuda_dev, T, Hq, Hk, D):
    from veomni_b200 import functional as F

    g = torch.Generator().manual_seed(T * 7 + Hq)
    q = torch.randn(T, Hq, D, generator=g).to(BF)
    k = torch.randn(T, Hk, D, generator=g).to(BF)
    wq = (1 + 0.2 * torch.randn(D, generator=g)).to(BF)
    wk = (1 + 0.2 * torch.randn(D, generator=g)).to(BF)
    cos, sin = o_ops.rotary_cos_sin(torch.arange(T)[None], D, 1e6, BF)
    cos, sin = cos[0], sin[0]
    args = [t.to(cuda_dev).requires_grad_(True) for t in (q, k, wq, wk)]
    qo, ko = F.qknorm_rope(*args, cos.to(cuda_dev), sin.to(cuda_dev), 1e-6)
    # oracle on bf16 tensors follows the reference's rounding points; compare loosely (rope rounding)
    qr, kr = o_ops.qknorm_rope(q, k, wq, wk, cos, sin, 1e-6)
    _close(qo, qr, atol=3e-2, rtol=1.6e-2, what="q")
    _close(ko, kr, atol=3e-2, rtol=1.6e-2, what="k")
    dq = torch.randn(qo.shape, generator=g).to(BF)
    dk = torch.randn(ko.shape, generator=g).to(BF)
    torch.autograd.backward([qo, ko], [dq.to(cuda_dev), dk.to(cuda_dev)])
    ref = [t.float().requires_grad_(True) for t in (q, k, wq, wk)]
    a, b = o_ops.qknorm_rope(*ref, cos.float(), sin.float(), 1e-6)
    torch.autograd.backward([a, b], [dq.float(), dk.float()])
    _close(args[0].grad, ref[0].grad, atol=3e-2, rtol=2e-2, what="dq")
    _close(args[1].grad, ref[1].grad, atol=3e-2, rtol=2e-2, what="dk")
    for i, nm in ((2, "dwq"), (3, "dwk")):
        s = max(1.0, float(ref[i].grad.abs().max()))
        _close(args[i].grad.float() / s, ref[i].grad / s, atol=2e-2, rtol=2e-2, what=nm)


def test_swiglu_matches_reference_fixture(cuda_dev, golden):
    from veomni_b200 import functional as F

    f = golden("ops.pt")["swiglu/bf16"]
    act = F.silu_mul(f["gate"].to(cuda_dev), f["up"].to(cuda_dev))
    _close(act, f["act"], atol=1e-2, rtol=8e-3)
    assert (act.cpu() == f["act"]).float().mean().item() > 0.98


@pytest.mark.parametrize("rows,cols", [(1, 8), (37, 160), (4096, 12288), (300, 768)])
def test_swiglu_fwd_bwd_vs_oracle(cuda_dev, rows, cols):
    from veomni_b200 import functional as F

    g = torch.Generator().manual_seed(rows + cols)
    gate = (2 * torch.randn(rows, cols, generator=g)).to(BF)
    up = torch.randn(rows, cols, generator=g).to(BF)
    d = torch.randn(rows, cols, generator=g).to(BF)
    gg, ug = gate.to(cuda_dev).requires_grad_(True), up.to(cuda_dev).requires_grad_(True)
    out = F.silu_mul(gg, ug)
    out.backward(d.to(cuda_dev))
    _close(out, o_ops.silu_mul(gate, up), atol=1e-2, rtol=8e-3, what="out")
    gf, uf = gate.float().requires_grad_(True), up.float().requires_grad_(True)
    o_ops.silu_mul(gf, uf).backward(d.float())
    _close(gg.grad, gf.grad, atol=2e-2, rtol=1e-2, what="dgate")
    _close(ug.grad, uf.grad, atol=2e-2, rtol=1e-2, what="dup")


def test_swiglu_merged_fc1_strided_views(cuda_dev):
    """The MoE path feeds the two halves of a merged [rows, 2I] fc1 output (moe_layer.py:339-346)."""
    from veomni_b200 import functional as F

    g = torch.Generator().manual_seed(0)
    fc1 = torch.randn(64, 2 * 96, generator=g).to(BF)
    a, b = fc1.to(cuda_dev).chunk(2, dim=-1)
    out = F.silu_mul(a, b)
    ra, rb = fc1.chunk(2, dim=-1)
    _close(out, o_ops.silu_mul(ra, rb), atol=1e-2, rtol=8e-3)


def test_multi_tensor_sumsq_and_scale(cuda_dev):
    """Gradient-clip kernels vs torch: ragged sizes (scalar heads/tails, multi-entry tensors), fp32 and bf16,
    determinism, and the no-op at coefficient 1."""
    from veomni_b200.clip_grad_norm import multi_scale_, multi_sumsq

    g = torch.Generator(device=cuda_dev).manual_seed(5)
    sizes = [1, 3, 4096, 4097, (1 << 20) + 5, 3 * (1 << 20), 12345]
    for dtype, rtol in ((torch.float32, 2e-6), (torch.bfloat16, 2e-6)):
        base = torch.randn(sum(sizes) + 3, generator=g, device=cuda_dev).to(dtype)
        ts, off = [], 1  # offset 1: misaligned starts
        for n in sizes:
            ts.append(base[off:off + n])
            off += n
        ref = sum((t.double() ** 2).sum() for t in ts)
        got = multi_sumsq(ts)
        assert abs(got.double().item() - ref.item()) / ref.item() < rtol
        assert torch.equal(got, multi_sumsq(ts))
        before = [t.clone() for t in ts]
        multi_scale_(ts, torch.ones((), device=cuda_dev))
        assert all(torch.equal(a, b) for a, b in zip(ts, before))
        multi_scale_(ts, torch.tensor(0.37, device=cuda_dev))
        for a, b in zip(ts, before):
            assert torch.equal(a, (b.float() * torch.tensor(0.37, device=cuda_dev)).to(dtype))


def test_clip_grad_norm_matches_torch(cuda_dev):
    """clip_grad_norm (dense path) against torch.nn.utils.clip_grad_norm_ on the same gradients: norm to 1e-6 rel.,
    clipped gradients bit-exact up to the coefficient's rounding (rtol 1e-6)."""
    from veomni_b200.clip_grad_norm import clip_grad_norm

    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(300, 257), torch.nn.Linear(257, 19)).to(cuda_dev)
    m2 = torch.nn.Sequential(torch.nn.Linear(300, 257), torch.nn.Linear(257, 19)).to(cuda_dev)
    m2.load_state_dict(m.state_dict())
    x = torch.randn(64, 300, device=cuda_dev)
    for mod in (m, m2):
        (mod(x) ** 2).sum().backward()
    total = clip_grad_norm(m, 0.5)
    ref = torch.nn.utils.clip_grad_norm_(m2.parameters(), 0.5)
    torch.testing.assert_close(total, ref, atol=0, rtol=1e-6)
    for a, b in zip(m.parameters(), m2.parameters()):
        torch.testing.assert_close(a.grad, b.grad, atol=1e-12, rtol=2e-6)


def test_b200_adamw_matches_torch_fused(cuda_dev):
    """B200AdamW against torch.optim.AdamW(fused=True): same hyper-parameters, 5 steps, ragged shapes (multi-entry tensors,
    unaligned views); parameters within 2e-6 relative (fp32, different but equivalent operation order)."""
    from veomni_b200.optim import B200AdamW

    torch.manual_seed(0)
    shapes = [(3,), (257, 129), (1 << 20,), ((1 << 20) + 7,), (4096, 1024)]
    ref = [torch.nn.Parameter(torch.randn(*s, device=cuda_dev)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    kw = dict(lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    o_ref, o_mine = torch.optim.AdamW(ref, fused=True, **kw), B200AdamW(mine, **kw)
    for step in range(5):
        for a, b in zip(ref, mine):
            g = torch.randn_like(a)
            a.grad, b.grad = g, g.clone()
        o_ref.step()
        o_mine.step()
    for a, b in zip(ref, mine):
        torch.testing.assert_close(b, a, atol=1e-7, rtol=2e-6)
    # grad_scale folds a clip coefficient into the step
    for a, b in zip(ref, mine):
        g = torch.randn_like(a)
        a.grad, b.grad = g * 0.25, g.clone()
    o_ref.step()
    o_mine.step(grad_scale=torch.tensor(0.25, device=cuda_dev))
    for a, b in zip(ref, mine):
        torch.testing.assert_close(b, a, atol=1e-7, rtol=2e-6)


def test_b200_adamw_master_weights_bf16_grads(cuda_dev):
    """master_weights=True (the world_size-1 path): bf16 model copy + bf16 gradients, fp32 master / moments inside the
    optimizer. Reference: torch.optim.AdamW(fused=True) on fp32 copies fed the same bf16 gradient values (bf16 -> fp32 is
    exact); the bf16 model parameter must be the fp32 master rounded to nearest-even, bit for bit."""
    from veomni_b200.optim import B200AdamW

    torch.manual_seed(1)
    shapes = [(5,), (257, 129), ((1 << 20) + 3,), (2048, 1024)]
    ref = [torch.nn.Parameter(torch.randn(*s, device=cuda_dev)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    kw = dict(lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    o_ref, o_mine = torch.optim.AdamW(ref, fused=True, **kw), B200AdamW(mine, master_weights=True, **kw)
    assert all(p.dtype == torch.bfloat16 for p in mine)
    for a, b in zip(ref, mine):
        assert torch.equal(b.data, a.data.to(torch.bfloat16))
    for step in range(4):
        coef = torch.tensor(0.5 if step == 3 else 1.0, device=cuda_dev)
        for a, b in zip(ref, mine):
            g = torch.randn_like(a).to(torch.bfloat16)
            a.grad, b.grad = g.float() * coef, g.clone()
        o_ref.step()
        o_mine.step(grad_scale=coef if step == 3 else None)
    for a, b in zip(ref, mine):
        master = o_mine.state[b]["master"]
        torch.testing.assert_close(master, a.data, atol=1e-7, rtol=2e-6)
        assert torch.equal(b.data, master.to(torch.bfloat16))
        assert o_mine.state[b]["exp_avg"].dtype == torch.float32


@pytest.mark.parametrize("rows,cols", [(1, 1024), (37, 2048), (4096, 4096), (65, 5120), (9, 8192)])
def test_fused_add_rms_norm_equals_add_then_norm(cuda_dev, rows, cols):
    """fused_add_rms_norm == (residual + x in bf16, then rms_norm): forward bit-exact (same rounding points), backward
    equal to autograd through the unfused pair up to the bf16 rounding of the gradient sum (1 ulp)."""
    from veomni_b200 import functional as F

    g = torch.Generator().manual_seed(rows + cols)
    x, r = (torch.randn(rows, cols, generator=g).to(BF).to(cuda_dev) for _ in range(2))
    w = (1 + 0.1 * torch.randn(cols, generator=g)).to(BF).to(cuda_dev)
    dy, dh = (torch.randn(rows, cols, generator=g).to(BF).to(cuda_dev) for _ in range(2))
    xa, ra, wa = (t.clone().requires_grad_(True) for t in (x, r, w))
    y1, h1 = F.fused_add_rms_norm(xa, ra, wa, 1e-6)
    (y1.float() * dy.float()).sum().add((h1.float() * dh.float()).sum()).backward()
    xb, rb, wb = (t.clone().requires_grad_(True) for t in (x, r, w))
    h2 = rb + xb
    y2 = F.rms_norm(h2, wb, 1e-6)
    (y2.float() * dy.float()).sum().add((h2.float() * dh.float()).sum()).backward()
    assert torch.equal(h1, h2) and torch.equal(y1, y2)
    _close(xa.grad, xb.grad, atol=2e-2, rtol=8e-3, what="dx")
    _close(ra.grad, rb.grad, atol=2e-2, rtol=8e-3, what="dresidual")
    _close(wa.grad, wb.grad, atol=1e-1, rtol=2e-2, what="dw")
