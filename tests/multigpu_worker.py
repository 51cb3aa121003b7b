"""Multi-GPU parity worker (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_worker.py [--bench]

Checks the NVLink pull collectives against the oracle's definition of each collective (rank-order
slicing) using NCCL only to move the *expected* data around, then FSDP2 end-to-end with the custom
comm installed vs PyTorch's default NCCL comm.  Prints ``WORKER OK`` on every rank when all pass.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle import comm as o_comm  # noqa: E402  (checker)


def gather_all(x):
    out = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(out, x.contiguous())
    return out


def test_barrier_and_allgather(symm, rank, world, dev):
    for it in range(3):
        symm.barrier()
    for numel in (8, 1000, 4096 * 3 + 2, 1 << 20):
        for dtype in (torch.bfloat16, torch.float32):
            out = symm.empty((numel * world,), dtype, arena="fsdp_ag")
            out.fill_(-1)
            g = torch.Generator(device="cpu").manual_seed(1000 * rank + numel)
            shard = torch.randn(numel, generator=g).to(dtype).to(dev)
            out[rank * numel : (rank + 1) * numel].copy_(shard)
            symm.all_gather_inplace(out, numel, 0)
            expect = torch.cat(gather_all(shard))
            assert torch.equal(out, expect), f"allgather mismatch numel={numel} dtype={dtype}"
    symm.check()


def test_reduce_scatter(symm, rank, world, dev):
    for chunk in (1, 7, 1024, 4099, 1 << 18):
        inp = symm.empty((chunk * world,), torch.float32, arena="fsdp_rs")
        g = torch.Generator(device="cpu").manual_seed(77 * rank + chunk)
        x = torch.randn(chunk * world, generator=g).to(dev)
        inp.copy_(x)
        out = torch.empty(chunk, dtype=torch.float32, device=dev)
        symm.reduce_scatter_f32(inp, out, 1.0 / world, 1)
        xs = gather_all(x)
        # oracle: fixed rank order 0..N-1 in fp32, then the divide (== multiply by 1/world for powers of two)
        ref = o_comm.fsdp_reduce_scatter([t.cpu() for t in xs], torch.float32, None)[rank] * (1.0 / world)
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), ref), f"reduce_scatter mismatch chunk={chunk}: {(out.cpu() - ref).abs().max()}"
        if rank == 0:
            print(f"  reduce_scatter chunk={chunk} ok", flush=True)
    symm.check()


def test_reduce_scatter_bf16(symm, rank, world, dev):
    """bf16-pull reduce-scatter == fp32 reduce-scatter of the fp32 copies (bit-exact) == the fixed-order oracle."""
    for chunk in (1, 7, 1024, 4099, 1 << 18):
        g = torch.Generator(device="cpu").manual_seed(91 * rank + chunk)
        xb = torch.randn(chunk * world, generator=g).to(torch.bfloat16).to(dev)
        # bf16 gradients packed into the first half of an fp32-sized buffer, as the patched copy-in leaves them
        buf = symm.empty((chunk * world,), torch.float32, arena="fsdp_rs")
        buf.view(torch.bfloat16)[: chunk * world].copy_(xb)
        out = torch.empty(chunk, dtype=torch.float32, device=dev)
        symm.reduce_scatter_bf16(buf, chunk, out, 1.0 / world, 1)
        inp = symm.empty((chunk * world,), torch.float32, arena="fsdp_rs")
        inp.copy_(xb.float())
        out32 = torch.empty(chunk, dtype=torch.float32, device=dev)
        symm.reduce_scatter_f32(inp, out32, 1.0 / world, 1)
        xs = gather_all(xb.float())
        ref = o_comm.fsdp_reduce_scatter([t.cpu() for t in xs], torch.float32, None)[rank] * (1.0 / world)
        torch.cuda.synchronize()
        assert torch.equal(out, out32), f"bf16-pull != fp32 reduce-scatter, chunk={chunk}"
        assert torch.equal(out.cpu(), ref), f"bf16-pull reduce_scatter vs oracle, chunk={chunk}"
    symm.check()


def test_reduce_scatter_push_oracle(symm, rank, world, dev):
    """Fused copy-in reduce-scatter == oracle: FSDP2's chunk_cat copy-in + reduce-scatter in fixed rank order + divide."""
    from veomni_b200.fsdp_comm import pack_plan

    shapes = [(64, 24), (9, 5), (33,), (128, 16)]
    g = torch.Generator(device="cpu").manual_seed(17 * rank + 1)
    grads = [torch.randn(*s, generator=g).to(torch.bfloat16).to(dev) for s in shapes]
    plan, row = pack_plan(shapes, world)
    staging = symm.empty((row * world,), torch.float32, arena="fsdp_rs")
    out = torch.empty(row, dtype=torch.float32, device=dev)
    desc = []
    for t, (numel, chunk, _off) in zip(grads, plan):
        desc += [t.data_ptr(), numel, chunk]
    symm.reduce_scatter_push_bf16(staging, desc, row, out, 1.0 / world, 1)
    # oracle input: every rank's fp32 chunk_cat buffer (what torch's copy-in produces), gathered with NCCL
    mine = torch.empty(world, row, dtype=torch.float32, device=dev)
    torch._chunk_cat([t.float() for t in grads], dim=0, num_chunks=world, out=mine)
    xs = gather_all(mine.flatten())
    ref = o_comm.fsdp_reduce_scatter([t.cpu() for t in xs], torch.float32, None)[rank] * (1.0 / world)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref), f"push reduce-scatter vs oracle: {(out.cpu() - ref).abs().max()}"
    symm.check()


def test_ulysses(rank, world, dev):
    from veomni_b200 import ulysses as U

    for (Sl, H, D) in ((6, 4 * world, 8), (128, 8 * world, 128), (33, 2 * world, 64)):
        g = torch.Generator(device="cpu").manual_seed(5 * rank + Sl)
        x = torch.randn(Sl, H, D, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
        y = U.gather_seq_scatter_heads(x, seq_dim=0, head_dim=1)
        xs = [t.cpu() for t in gather_all(x.detach())]
        ref = o_comm.gather_seq_scatter_heads(xs, seq_dim=0, head_dim=1)[rank]
        assert torch.equal(y.detach().cpu(), ref), f"ulysses gather_seq_scatter_heads mismatch {Sl, H, D}"
        z = U.gather_heads_scatter_seq(y, head_dim=1, seq_dim=0)
        assert torch.equal(z.detach(), x.detach()), "ulysses round trip"
        # autograd: d/dx of sum(w * y) is the reverse exchange of w
        w = torch.randn(y.shape, generator=torch.Generator().manual_seed(9 + rank)).to(torch.bfloat16).to(dev)
        (y * w).sum().backward()
        ws = [t.cpu() for t in gather_all(w)]
        gref = o_comm.gather_heads_scatter_seq(ws, head_dim=1, seq_dim=0)[rank]
        assert torch.equal(x.grad.cpu(), gref), "ulysses backward"
        # 4-D [1, S, H, D] layout and the fused q/k/v launch
        q = x.detach()[None]
        k = x.detach()[None, :, : 2 * world].contiguous()
        qo, ko, vo = U.gather_seq_scatter_heads_qkv(q, k, k, seq_dim=1, head_dim=2)
        assert torch.equal(qo[0].cpu(), ref)
        ks = [t.cpu() for t in gather_all(k[0])]
        assert torch.equal(ko[0].cpu(), o_comm.gather_seq_scatter_heads(ks, 0, 1)[rank])
        assert torch.equal(vo, ko)


def test_fsdp(rank, world, dev):
    from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

    from veomni_b200.fsdp_comm import install_fsdp_comm

    def build():
        torch.manual_seed(3)
        m = torch.nn.Sequential(*[torch.nn.Linear(256, 256, bias=False) for _ in range(4)]).to(dev)
        mpp = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32)
        for layer in m:
            fully_shard(layer, mp_policy=mpp)
        fully_shard(m, mp_policy=mpp)
        return m

    g = torch.Generator(device="cpu").manual_seed(40 + rank)
    x = torch.randn(32, 256, generator=g).to(dev)
    ref_m = build()
    ref_m(x).float().square().mean().backward()
    ref = {n: p.grad.to_local().clone() for n, p in ref_m.named_parameters()}
    got = {}
    for pack in (True, False):  # bf16-packed copy-in + bf16 pull, and torch's fp32 copy-in + fp32 pull
        m = build()
        install_fsdp_comm(m, pack_bf16=pack)
        for step in range(2):  # second step exercises buffer reuse
            for p in m.parameters():
                p.grad = None
            out = m(x)
            out.float().square().mean().backward()
        torch.cuda.synchronize()
        got[pack] = {n: p.grad.to_local().clone() for n, p in m.named_parameters()}
        for n, g in got[pack].items():
            # same bf16 GEMMs; only the fp32 summation order of the reduce-scatter differs from NCCL's ring
            torch.testing.assert_close(g, ref[n], atol=1e-6, rtol=1e-4, msg=lambda s, n=n: f"fsdp grad {n} (pack={pack}): {s}")
    for n in got[True]:
        assert torch.equal(got[True][n], got[False][n]), f"bf16-packed reduce-scatter changed the gradient of {n}"


def test_async_ulysses(rank, world, dev):
    """Async-Ulysses q/k/v and output projections == the synchronous path (projection -> exchange -> q/k RMSNorm), the parity
    target of the reference's own test (tests/parallel/ulysses/test_async_ulysses.py:112-115): outputs bit-exact (same kernels,
    same order), gradients to GEMM-summation-order tolerance."""
    from veomni_b200 import functional as F
    from veomni_b200 import ulysses as U
    from veomni_b200.async_ulysses import async_ulysses_output_projection, async_ulysses_qkv_projection

    BF = torch.bfloat16
    H, nq, nkv, D = 256, 4 * world, max(2, world), 64
    Sl = 40 + 0  # local tokens per rank (same on every rank)
    g = torch.Generator().manual_seed(900)  # same weights everywhere
    qw = (0.05 * torch.randn(nq * D, H, generator=g)).to(BF).to(dev)
    kw = (0.05 * torch.randn(nkv * D, H, generator=g)).to(BF).to(dev)
    vw = (0.05 * torch.randn(nkv * D, H, generator=g)).to(BF).to(dev)
    ow = (0.05 * torch.randn(H, nq * D, generator=g)).to(BF).to(dev)
    nqw = (1 + 0.1 * torch.randn(D, generator=g)).to(BF).to(dev)
    nkw = (1 + 0.1 * torch.randn(D, generator=g)).to(BF).to(dev)
    g2 = torch.Generator().manual_seed(901 + rank)
    hs = torch.randn(1, Sl, H, generator=g2).to(BF).to(dev)
    res = {}
    for mode in ("sync", "async"):
        leaves = [t.clone().requires_grad_(True) for t in (hs, qw, kw, vw, ow, nqw, nkw)]
        h, a, b, c, o, n1, n2 = leaves
        if mode == "async":
            q, k, v = async_ulysses_qkv_projection(h, 1, 2, a, None, b, None, c, None, "rmsnorm", n1, None, n2, None, D, 1e-6,
                                                   Sl * world, D)
        else:
            q = U.gather_seq_scatter_heads(torch.nn.functional.linear(h, a).view(1, -1, nq, D), seq_dim=1, head_dim=2)
            k = U.gather_seq_scatter_heads(torch.nn.functional.linear(h, b).view(1, -1, nkv, D), seq_dim=1, head_dim=2)
            v = U.gather_seq_scatter_heads(torch.nn.functional.linear(h, c).view(1, -1, nkv, D), seq_dim=1, head_dim=2)
            q, k = F.rms_norm(q, n1, 1e-6), F.rms_norm(k, n2, 1e-6)
        att = q * 0.5 + k.repeat_interleave(nq // nkv, dim=2) * 0.25 + v.repeat_interleave(nq // nkv, dim=2)  # stand-in for attention
        if mode == "async":
            out = async_ulysses_output_projection(att, 1, 2, o, None, att.shape[1])
        else:
            out = torch.nn.functional.linear(U.gather_heads_scatter_seq(att, head_dim=2, seq_dim=1).reshape(1, Sl, -1), o)
        out.float().square().mean().backward()
        torch.cuda.synchronize()
        res[mode] = ([q.detach(), k.detach(), v.detach(), out.detach()], [t.grad for t in leaves])
    for x, y in zip(res["sync"][0], res["async"][0]):
        assert torch.equal(x, y), "async-Ulysses forward differs from the synchronous path"
    for i, (x, y) in enumerate(zip(res["sync"][1], res["async"][1])):
        sc = max(1e-6, float(x.float().abs().max()))
        torch.testing.assert_close(y.float() / sc, x.float() / sc, atol=2e-2, rtol=2e-2, msg=lambda m, i=i: f"async-Ulysses grad {i}: {m}")


def test_ulysses_model(rank, world, dev):
    """Toy Qwen3 through the host caller with Ulysses SP over all ranks vs the reference fixture (loss, grad-norm)."""
    from veomni_b200.host_qwen3 import Qwen3Config, Qwen3ForCausalLM

    f = torch.load(REPO / "tests" / "golden" / "qwen3_toy.pt", weights_only=False)
    cfg = Qwen3Config.from_hf_dict(f["config"])
    if cfg.num_attention_heads % world or sum(f["seq_lens"]) % world:
        return
    model = Qwen3ForCausalLM(cfg)
    model.load_state_dict(f["state_dict"])
    model = model.to(dev).to(torch.bfloat16)
    model.sp_group = dist.group.WORLD
    model.train()
    lens = f["seq_lens"]
    T = sum(lens)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    # SequenceParallelCollator (veomni/data/data_collator.py:336-389): shift labels on the full row, then slice
    shift = torch.nn.functional.pad(f["labels"], (0, 1), value=-100)[..., 1:]
    sl = slice(rank * T // world, (rank + 1) * T // world)
    ids, pos, sh = f["input_ids"][:, sl].to(dev), f["position_ids"][:, sl].to(dev), shift[:, sl].to(dev)
    loss_local = model(ids, pos, cu, max(lens), shift_labels=sh)
    n_local = (sh != -100).sum().float()
    n_tot = n_local.clone()
    dist.all_reduce(n_tot)
    (loss_local * n_local / n_tot).backward()
    loss = (loss_local.detach().float() * n_local / n_tot)
    dist.all_reduce(loss)
    ref = float(f["loss"])
    assert abs(float(loss) - ref) / ref < 2e-2, (float(loss), ref)
    sq = torch.zeros((), device=dev)
    for p_ in model.parameters():
        g = p_.grad.float()
        dist.all_reduce(g)  # every rank holds the full weights: SP grads add up
        sq += (g * g).sum()
    gn = float(sq.sqrt())
    assert abs(gn - float(f["grad_norm"])) / float(f["grad_norm"]) < 5e-2, (gn, float(f["grad_norm"]))


def test_ep_model(rank, world, dev):
    """Toy Qwen3-MoE through the host caller: EP over all ranks (ParallelPlan slicing + NVLink dispatch/combine)
    vs the same model without EP on the same tokens — loss and hidden-state gradients must agree."""
    from veomni_b200 import moe as M
    from veomni_b200.ep import EPContext
    from veomni_b200.host_qwen3_moe import Qwen3MoeConfig, Qwen3MoeForCausalLM

    cfg = Qwen3MoeConfig(vocab_size=256, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=64, num_experts=4 * world, num_experts_per_tok=2,
                         moe_intermediate_size=128)
    torch.manual_seed(0)
    ref = Qwen3MoeForCausalLM(cfg).to(dev)
    ref.init_weights(seed=0)
    ref = ref.to(torch.bfloat16)
    ep_model = Qwen3MoeForCausalLM(cfg).to(dev).to(torch.bfloat16)
    ep_model.load_state_dict(ref.state_dict())
    ep_model.get_parallel_plan().apply(ep_model, ep_size=world, ep_rank=rank)
    assert ep_model.model.layers[0].mlp.experts.gate_up_proj.shape[0] == 4
    g = torch.Generator().manual_seed(50 + rank)
    lens = [70, 58]
    ids = torch.randint(0, 256, (1, sum(lens)), generator=g).to(dev)
    pos = torch.cat([torch.arange(n) for n in lens])[None].to(dev)
    cu = torch.tensor([0, 70, 128], dtype=torch.int32, device=dev)
    ref.train(); ep_model.train()
    M.set_ep_group(None)
    loss_ref = ref(ids, pos, cu, 70, labels=ids)
    loss_ref.backward()
    M.set_ep_group(EPContext())
    try:
        loss_ep = ep_model(ids, pos, cu, 70, labels=ids)
        loss_ep.backward()
    finally:
        M.set_ep_group(None)
    torch.testing.assert_close(loss_ep.float(), loss_ref.float(), atol=2e-2, rtol=2e-2)
    ga, gb = ep_model.model.embed_tokens.weight.grad.float(), ref.model.embed_tokens.weight.grad.float()
    s = max(1e-6, float(gb.abs().max()))
    torch.testing.assert_close(ga / s, gb / s, atol=5e-2, rtol=5e-2)
    # expert weight grads: the EP rank's slice equals the sum over ranks of the non-EP grads for those experts
    full = ref.model.layers[0].mlp.experts.down_proj.grad.float().clone()
    dist.all_reduce(full)
    mine = ep_model.model.layers[0].mlp.experts.down_proj.grad.float()
    s = max(1e-6, float(full.abs().max()))
    torch.testing.assert_close(mine / s, full[rank * 4 : (rank + 1) * 4] / s, atol=5e-2, rtol=5e-2)


def test_ep(rank, world, dev):
    """EP dispatch/combine + expert MLP vs the oracle (and vs the reference run stored in tests/golden/multirank.pt)."""
    from oracle import moe as o_moe
    from veomni_b200 import ep as EP

    ctx = EP.EPContext()
    BF = torch.bfloat16
    if world == 2:  # the fixture was produced by the reference on a 2-rank gloo group
        fx = torch.load(REPO / "tests" / "golden" / "multirank.pt", weights_only=False)["ranks"]
        me = fx[rank]["ep"]
        E = 8
        hs, idx, rw = me["hs"].to(BF).to(dev), me["idx"].to(dev), me["rw"].to(dev)
        tokens, plan = EP.ep_dispatch(ctx, hs, idx, E)
        assert plan.input_splits == me["input_splits"] and plan.output_splits == me["output_splits"], "EP split sizes"
        assert torch.equal(tokens.cpu(), me["tokens"].to(BF)), "EP dispatched tokens differ from the reference"
        final = EP.ep_combine(tokens * 2.0, rw, plan)
        torch.testing.assert_close(final.float().cpu(), me["final"], atol=3e-2, rtol=3e-2)
    # larger random case with real expert MLPs, forward + backward
    E, K, H, I = 8 * world, 4, 256, 128
    T = 300 + 17 * rank
    g = torch.Generator().manual_seed(1000 + rank)
    hs = (0.5 * torch.randn(T, H, generator=g)).to(BF)
    logits = torch.randn(T, E, generator=g)
    if rank == 0:
        logits[:, 1] = -1e9  # nobody on rank 0 picks expert 1
    rw, idx = torch.topk(torch.softmax(logits, -1), K, dim=-1)
    rw = (rw / rw.sum(-1, keepdim=True)).to(BF)
    gw = torch.Generator().manual_seed(7)  # same weights on every rank
    w1 = (0.1 * torch.randn(E, 2 * I, H, generator=gw)).to(BF)
    w2 = (0.1 * torch.randn(E, H, I, generator=gw)).to(BF)
    dy = (0.1 * torch.randn(T, H, generator=g)).to(BF)
    el = E // world
    hs_d = hs.to(dev).requires_grad_(True)
    rw_d = rw.to(dev).requires_grad_(True)
    w1_d = w1[rank * el : (rank + 1) * el].to(dev).requires_grad_(True)
    w2_d = w2[rank * el : (rank + 1) * el].to(dev).requires_grad_(True)
    out = EP.ep_fused_moe_forward(ctx, E, rw_d, idx.to(dev), hs_d, w1_d, w2_d)
    out.backward(dy.to(dev))
    torch.cuda.synchronize()

    def gather_var(x):  # variable T per rank
        lens = [300 + 17 * r for r in range(world)]
        res = []
        for r in range(world):
            buf = torch.empty((lens[r],) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
            if r == rank:
                buf.copy_(x)
            dist.broadcast(buf, src=r)
            res.append(buf.cpu())
        return res

    all_hs, all_rw, all_idx, all_dy = gather_var(hs.to(dev)), gather_var(rw.to(dev)), gather_var(idx.to(dev)), gather_var(dy.to(dev))
    # oracle on fp32 copies of the bf16 inputs, all ranks at once
    hs_f = [t.float().requires_grad_(True) for t in all_hs]
    rw_f = [t.float().requires_grad_(True) for t in all_rw]
    w1_f, w2_f = w1.float().requires_grad_(True), w2.float().requires_grad_(True)
    outs, disp = o_moe.ep_moe_forward(hs_f, rw_f, all_idx, E, w1_f, w2_f)
    assert plan_equal(EP, ctx, idx.to(dev), E, H, disp[rank])
    torch.testing.assert_close(out.float().cpu(), outs[rank].detach(), atol=2e-2, rtol=3e-2)
    torch.autograd.backward(outs, [d.float() for d in all_dy])
    for name, got, want in (("d_hidden", hs_d.grad, hs_f[rank].grad), ("d_routing", rw_d.grad, rw_f[rank].grad),
                            ("d_fc1", w1_d.grad, w1_f.grad[rank * el : (rank + 1) * el]),
                            ("d_fc2", w2_d.grad, w2_f.grad[rank * el : (rank + 1) * el])):
        sc = max(1e-6, float(want.abs().max()))
        torch.testing.assert_close(got.float().cpu() / sc, want / sc, atol=5e-2, rtol=5e-2, msg=lambda m, n=name: f"EP {n}: {m}")


def plan_equal(EP, ctx, idx, E, H, ref):
    plan = EP.make_plan(ctx, idx, E, H)
    ok = plan.input_splits == ref["input_splits"] and plan.output_splits == ref["output_splits"]
    ok = ok and torch.equal(plan.cumsum_local.cpu().long(), ref["cumsum"].long())
    return ok


def bench(symm, rank, world, dev):
    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        symm.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    res = []
    U = 192_950_000 // world * world  # Qwen3-8B layer unit
    out = symm.empty((U,), torch.bfloat16, arena="fsdp_ag")
    for ctas in (16, 32, 64):
        ms = timeit(lambda: symm.all_gather_inplace(out, U // world, 0, ctas))
        nbytes = (world - 1) / world * U * 2
        res.append({"kernel": f"fsdp_allgather[U=193M bf16,N={world},ctas={ctas}]", "ms": round(ms, 3), "nvlink_in_GBps": round(nbytes / ms / 1e6, 1)})
    # the fused copy-out variant: Qwen3-8B layer parameter shapes, destinations in ordinary device memory
    H, I, Hq, Hk, D = 4096, 12288, 32, 8, 128
    shapes = [(Hq * D, H), (Hk * D, H), (Hk * D, H), (H, Hq * D), (D,), (D,), (I, H), (I, H), (H, I), (H,), (H,)]
    numels = [(s_[0] + world - 1) // world * (s_[1] if len(s_) > 1 else 1) for s_ in shapes]
    row = sum(numels)
    buf = symm.empty((row * world,), torch.bfloat16, arena="fsdp_ag")
    dsts = [torch.empty(n * world, dtype=torch.bfloat16, device=dev) for n in numels]
    table, off = [], 0
    for n, d in zip(numels, dsts):
        table += [off * 2, n * 2, d.data_ptr()]
        off += n
    for ctas in (8, 16, 24, 32, 48, 64):
        ms = timeit(lambda: symm.all_gather_scatter(buf, row, table, 0, ctas))
        res.append({"kernel": f"fsdp_allgather+copy-out[Qwen3-8B layer unit,N={world},ctas={ctas}]", "ms": round(ms, 3),
                    "nvlink_in_GBps": round((world - 1) * row * 2 / ms / 1e6, 1)})
    del buf, dsts
    nc = torch.empty(U, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: dist.all_gather_into_tensor(nc, nc[rank * (U // world) : (rank + 1) * (U // world)]))
    res.append({"kernel": f"(lib) nccl all_gather[U=193M bf16,N={world}]", "ms": round(ms, 3), "nvlink_in_GBps": round((world - 1) / world * U * 2 / ms / 1e6, 1)})
    del out
    inp = symm.empty((U,), torch.float32, arena="fsdp_rs")
    o = torch.empty(U // world, dtype=torch.float32, device=dev)
    for ctas in (16, 32, 64):
        ms = timeit(lambda: symm.reduce_scatter_f32(inp, o, 1.0 / world, 1, ctas))
        res.append({"kernel": f"fsdp_reducescatter[U=193M fp32,N={world},ctas={ctas}]", "ms": round(ms, 3), "nvlink_in_GBps": round((world - 1) / world * U * 4 / ms / 1e6, 1)})
    for ctas in (16, 32, 64):
        ms = timeit(lambda: symm.reduce_scatter_bf16(inp, U // world, o, 1.0 / world, 1, ctas))
        res.append({"kernel": f"fsdp_reducescatter_bf16pull[U=193M,N={world},ctas={ctas}]", "ms": round(ms, 3), "nvlink_in_GBps": round((world - 1) / world * U * 2 / ms / 1e6, 1)})
    # the fused copy-in variant: gradients read in place (bf16), pushed, reduced
    from veomni_b200.fsdp_comm import pack_plan

    grads = [torch.randn(*s_, device=dev, dtype=torch.bfloat16) for s_ in shapes]
    plan, prow = pack_plan(shapes, world)
    desc = []
    for t, (numel, chunk, _o) in zip(grads, plan):
        desc += [t.data_ptr(), numel, chunk]
    o2 = torch.empty(prow, dtype=torch.float32, device=dev)
    for ctas in (8, 16, 24, 32, 48, 64):
        ms = timeit(lambda: symm.reduce_scatter_push_bf16(inp, desc, prow, o2, 1.0 / world, 1, ctas))
        res.append({"kernel": f"fsdp_reducescatter push+reduce, copy-in fused[Qwen3-8B layer unit,N={world},ctas={ctas}]", "ms": round(ms, 3),
                    "nvlink_out_GBps": round((world - 1) * prow * 2 / ms / 1e6, 1)})
    del grads
    nci = torch.empty(U, dtype=torch.float32, device=dev)
    ms = timeit(lambda: dist.reduce_scatter_tensor(o, nci, op=dist.ReduceOp.AVG))
    res.append({"kernel": f"(lib) nccl reduce_scatter[U=193M fp32,N={world}]", "ms": round(ms, 3), "nvlink_in_GBps": round((world - 1) / world * U * 4 / ms / 1e6, 1)})
    del inp
    from veomni_b200 import ulysses as Uly

    S, Hq, Hk, D = 32768, 32, 8, 128
    if Hk % world == 0:
        q = torch.randn(S // world, Hq, D, device=dev, dtype=torch.bfloat16)
        k = torch.randn(S // world, Hk, D, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: Uly.all_to_all_many([q, k, k], 1, 0))
        nbytes = (world - 1) / world * (q.numel() + 2 * k.numel()) * 2
        res.append({"kernel": f"ulysses_a2a_qkv[S=32k,P={world}]", "ms": round(ms, 3), "nvlink_in_GBps": round(nbytes / ms / 1e6, 1)})
    if rank == 0:
        for r in res:
            print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from veomni_b200.symm import get_symmetric_memory

    def stage(name, fn, *fa):
        try:
            fn(*fa)
            torch.cuda.synchronize()
            symm_ref[0] and symm_ref[0].check()
            print(f"[rank {rank}] stage ok: {name}", flush=True)
        except Exception:
            import traceback

            print(f"[rank {rank}] stage FAILED: {name}", flush=True)
            traceback.print_exc()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(3)

    symm_ref = [None]

    def make():
        symm_ref[0] = get_symmetric_memory(None, (3 << 30) if a.bench else (256 << 20),
                                           {"fsdp_ag": 0.3, "fsdp_rs": 0.5, "misc": 0.2})

    stage("symmetric memory + IPC mapping", make)
    symm = symm_ref[0]
    stage("barrier x3", lambda: [symm.barrier() for _ in range(3)])
    stage("allgather", test_barrier_and_allgather, symm, rank, world, dev)
    stage("reduce_scatter", test_reduce_scatter, symm, rank, world, dev)
    stage("reduce_scatter bf16 pull", test_reduce_scatter_bf16, symm, rank, world, dev)
    # the world-size-generic kernel instantiation (what 8 GPUs run), forced on this group
    os.environ["VB200_RS_GENERIC"] = "1"
    stage("reduce_scatter (generic instantiation)", test_reduce_scatter, symm, rank, world, dev)
    stage("reduce_scatter bf16 pull (generic instantiation)", test_reduce_scatter_bf16, symm, rank, world, dev)
    os.environ.pop("VB200_RS_GENERIC")
    from veomni_b200 import selfcheck

    stage("all-gather with fused copy-out (oracle-free self-check)", selfcheck.check_allgather_scatter, symm, dev)
    stage("reduce-scatter with fused copy-in (oracle-free self-check)", selfcheck.check_reduce_scatter_push, symm, dev)
    stage("reduce-scatter with fused copy-in vs the fixed-order oracle", test_reduce_scatter_push_oracle, symm, rank, world, dev)
    os.environ["VB200_RS_GENERIC"] = "1"
    stage("reduce-scatter with fused copy-in (generic instantiation)", selfcheck.check_reduce_scatter_push, symm, dev)
    os.environ.pop("VB200_RS_GENERIC")
    stage("ulysses", test_ulysses, rank, world, dev)
    stage("all_to_all_images: uneven row exchange, forward + backward (self-check)", selfcheck.check_images, symm, dev)
    stage("async ulysses projections == synchronous path", test_async_ulysses, rank, world, dev)
    stage("fsdp2 custom comm", test_fsdp, rank, world, dev)
    stage("fsdp2 custom comm: every mode, ragged shapes, divide factor (self-check)", selfcheck.check_fsdp, dev, world)
    stage("EP dispatch/combine (self-check)", selfcheck.check_ep_dispatch, symm, dev)
    stage("expert parallel dispatch/combine", test_ep, rank, world, dev)
    stage("ulysses SP through the model (reference fixture)", test_ulysses_model, rank, world, dev)
    stage("expert parallel through the Qwen3-MoE caller", test_ep_model, rank, world, dev)
    if a.bench:
        stage("bench", bench, symm, rank, world, dev)
    dist.barrier()
    print(f"WORKER OK rank {rank}/{world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
