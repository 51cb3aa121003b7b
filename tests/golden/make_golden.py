"""Generate the golden fixtures in this directory from the *reference itself*.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONPATH=/root/reference:/root/repo python tests/golden/make_golden.py

For every piece of the hot path that the reference can execute on CPU it
  1. runs the reference function on seeded inputs,
  2. runs the oracle restatement (``oracle/``) on the same inputs and asserts agreement
     (bit-exact unless a tolerance is stated), which *pins* the oracle,
  3. stores inputs + reference outputs in ``*.pt`` fixtures for tests/test_oracle_golden.py and
     for the GPU parity tests.

Multi-rank pieces (Ulysses all-to-all, EP dispatch/combine, FSDP2) run the reference on a
2-process gloo group, exactly as SURVEY.md §8(c) describes (with the two documented stubs).
"""

from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, "/root/reference")

from oracle import attention as o_attn  # noqa: E402
from oracle import comm as o_comm  # noqa: E402
from oracle import loss as o_loss  # noqa: E402
from oracle import moe as o_moe  # noqa: E402
from oracle import ops as o_ops  # noqa: E402


def _stub_triton_group_gemm():
    """veomni.distributed.moe imports the Triton GroupGEMM at import time."""
    name = "veomni.ops.kernels.moe._kernels.kernel.group_gemm"
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.group_gemm_same_nk = None
        m.group_gemm_same_mn = None
        sys.modules[name] = m


def eq(a, b, what, atol=0.0, rtol=0.0):
    if atol == 0.0 and rtol == 0.0:
        assert torch.equal(a, b), f"{what}: oracle != reference (max diff {(a.float() - b.float()).abs().max()})"
    else:
        torch.testing.assert_close(a, b, atol=atol, rtol=rtol, msg=lambda m: f"{what}: {m}")
    print(f"  pinned: {what}")


# ----------------------------------------------------------------------------------------------
def gen_ops():
    from transformers import Qwen3Config
    from veomni.models.transformers.qwen3.generated import patched_modeling_qwen3_gpu as M

    out = {}
    g = torch.Generator().manual_seed(0)
    for dtype, tag in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
        # RMSNorm: hidden-size rows and head_dim rows
        for name, rows, cols in (("hidden", 12, 512), ("head", 40, 128)):
            x = (torch.randn(rows, cols, generator=g) * 1.5).to(dtype)
            norm = M.Qwen3RMSNorm(cols, eps=1e-6).to(dtype)
            with torch.no_grad():
                norm.weight.copy_((1 + 0.1 * torch.randn(cols, generator=g)).to(dtype))
            y_ref = norm(x)
            eq(o_ops.rms_norm(x, norm.weight.detach(), 1e-6), y_ref.detach(), f"rms_norm {name} {tag}")
            out[f"rms_norm/{name}/{tag}"] = {"x": x, "w": norm.weight.detach().clone(), "eps": 1e-6, "y": y_ref.detach()}

        # RoPE incl. the reference's cos/sin generation
        cfg = Qwen3Config(hidden_size=256, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                          max_position_embeddings=4096, rope_theta=1000000.0)
        rot = M.Qwen3RotaryEmbedding(cfg)
        S = 24
        pos = torch.cat([torch.arange(10), torch.arange(14)])[None]  # two packed sequences
        q = torch.randn(1, 4, S, 64, generator=g).to(dtype)
        k = torch.randn(1, 2, S, 64, generator=g).to(dtype)
        cos, sin = rot(q, pos)
        rope_theta = cfg.rope_parameters["rope_theta"] if hasattr(cfg, "rope_parameters") else cfg.rope_theta
        c2, s2 = o_ops.rotary_cos_sin(pos, 64, rope_theta, dtype)
        eq(c2, cos, f"rotary cos {tag}")
        eq(s2, sin, f"rotary sin {tag}")
        qe, ke = M.apply_rotary_pos_emb(q, k, cos, sin)
        qo, ko = o_ops.apply_rotary_pos_emb(q, k, cos, sin)
        eq(qo, qe, f"rope q {tag}")
        eq(ko, ke, f"rope k {tag}")
        out[f"rope/{tag}"] = {"q": q, "k": k, "cos": cos, "sin": sin, "pos": pos, "theta": rope_theta, "q_out": qe, "k_out": ke}

        # SwiGLU MLP
        cfg2 = Qwen3Config(hidden_size=64, intermediate_size=160)
        mlp = M.Qwen3MLP(cfg2).to(dtype)
        x = torch.randn(9, 64, generator=g).to(dtype)
        y_ref = mlp(x).detach()
        y_or = o_ops.swiglu_mlp(x, mlp.gate_proj.weight.detach(), mlp.up_proj.weight.detach(), mlp.down_proj.weight.detach())
        eq(y_or, y_ref, f"swiglu_mlp {tag}")
        gate, up = mlp.gate_proj(x).detach(), mlp.up_proj(x).detach()
        out[f"swiglu/{tag}"] = {"gate": gate, "up": up, "act": (mlp.act_fn(gate) * up).detach(),
                                "x": x, "wg": mlp.gate_proj.weight.detach().clone(),
                                "wu": mlp.up_proj.weight.detach().clone(), "wd": mlp.down_proj.weight.detach().clone(),
                                "y": y_ref}

    # Attention: the reference's eager attention with a block-diagonal causal mask == varlen packed
    cu = torch.tensor([0, 10, 24, 31], dtype=torch.int32)
    T, Hq, Hk, D = 31, 4, 2, 64
    q = torch.randn(1, Hq, T, D, generator=g)
    k = torch.randn(1, Hk, T, D, generator=g)
    v = torch.randn(1, Hk, T, D, generator=g)
    mask = torch.full((T, T), float("-inf"))
    for a, b in zip(cu[:-1].tolist(), cu[1:].tolist()):
        mask[a:b, a:b] = torch.triu(torch.full((b - a, b - a), float("-inf")), diagonal=1)
    mod = types.SimpleNamespace(num_key_value_groups=Hq // Hk, training=False)
    o_ref, _ = M.eager_attention_forward(mod, q, k, v, mask[None, None], scaling=D**-0.5)
    o_or, lse = o_attn.varlen_causal_attention(q[0].transpose(0, 1), k[0].transpose(0, 1), v[0].transpose(0, 1), cu)
    eq(o_or, o_ref[0], "varlen attention fp32", atol=2e-6, rtol=1e-5)
    out["attention/fp32"] = {"q": q[0].transpose(0, 1).contiguous(), "k": k[0].transpose(0, 1).contiguous(),
                             "v": v[0].transpose(0, 1).contiguous(), "cu": cu, "out": o_ref[0].contiguous(), "lse": lse}
    torch.save(out, HERE / "ops.pt")


# ----------------------------------------------------------------------------------------------
def gen_moe_local():
    """Pieces of the MoE path the reference can run in one process on CPU."""
    _stub_triton_group_gemm()
    from transformers import Qwen3MoeConfig
    from veomni.distributed.moe import moe_utils as U
    from veomni.models.transformers.qwen3_moe.generated import patched_modeling_qwen3_moe_gpu as MM

    out = {}
    g = torch.Generator().manual_seed(1)
    T, E, K, H, I = 48, 8, 2, 64, 32
    cfg = Qwen3MoeConfig(hidden_size=H, moe_intermediate_size=I, num_experts=E, num_experts_per_tok=K, norm_topk_prob=True)
    experts = MM.Qwen3MoeExperts(cfg)
    router = MM.Qwen3MoeTopKRouter(cfg)
    with torch.no_grad():
        experts.gate_up_proj.copy_(0.1 * torch.randn(E, 2 * I, H, generator=g))
        experts.down_proj.copy_(0.1 * torch.randn(E, H, I, generator=g))
        router.weight.copy_(torch.randn(E, H, generator=g))
    hs = 0.5 * torch.randn(T, H, generator=g)
    _, rw, idx = router(hs)
    # leave expert 5 empty on purpose (ragged M == 0 edge case)
    idx = torch.where(idx == 5, torch.full_like(idx, 6), idx)
    y_ref = experts(hs, idx, rw).detach()
    gu, dn = experts.gate_up_proj.detach(), experts.down_proj.detach()
    eq(o_moe.eager_moe_forward(E, rw, idx, hs, gu, dn), y_ref, "eager MoE experts fp32")
    y_fused, inter = o_moe.fused_moe_forward(E, rw, idx, hs, gu, dn)
    eq(y_fused, y_ref, "fused-order MoE vs eager fp32", atol=1e-5, rtol=1e-4)
    out["moe/fp32"] = {"hs": hs, "rw": rw.detach(), "idx": idx, "gate_up": gu.clone(), "down": dn.clone(), "y": y_ref,
                       "scatter_index": inter["scatter_index"], "splits": inter["splits"]}

    # reference moe_utils on CPU: permute / unpermute / weights idx / sort_chunks
    mask = torch.nn.functional.one_hot(idx, num_classes=E).permute(2, 1, 0)
    routing_map = mask.sum(dim=1)
    p_ref, m_ref = U.permute(hs, routing_map)
    p_or, m_or = o_moe.permute(hs, routing_map)
    eq(p_or, p_ref, "permute tokens")
    eq(m_or, m_ref, "permute mapping")
    w_ref = U.generate_weights_idx(rw.detach(), idx, E)
    eq(o_moe.generate_weights_idx(rw.detach(), idx, E), w_ref, "generate_weights_idx")
    u_ref = U.unpermute(p_ref, w_ref, hs.shape, m_ref, routing_map)
    eq(o_moe.unpermute(p_ref, w_ref, hs.shape, m_ref, routing_map), u_ref, "unpermute")
    out["moe_utils/fp32"] = {"routing_map": routing_map, "perm": p_ref, "mapping": m_ref, "weights_idx": w_ref, "unperm": u_ref}
    torch.save(out, HERE / "moe.pt")


# ----------------------------------------------------------------------------------------------
def _init_gloo(rank, world, path):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    store = dist.FileStore(path, world)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)


def _worker(rank, world, path, outdir):
    torch.set_num_threads(2)
    _init_gloo(rank, world, path)
    _stub_triton_group_gemm()
    res = {}
    group = dist.group.WORLD

    # ---- Ulysses packed [S, H, D] path (reference) -------------------------------------------
    from veomni.distributed.sequence_parallel import ulysses as UL

    g = torch.Generator().manual_seed(100 + rank)
    S_local, Hh, D = 6, 4, 8
    x = torch.randn(S_local, Hh, D, generator=g)
    y = UL.gather_seq_scatter_heads(x, seq_dim=0, head_dim=1, group=group)
    z = UL.gather_heads_scatter_seq(y, head_dim=1, seq_dim=0, group=group)
    assert torch.equal(z, x)
    res["ulysses"] = {"x": x, "gathered": y}

    # ---- EP dispatch / combine (reference) ---------------------------------------------------
    orig_ag = dist.all_gather_into_tensor

    def flat_ag(o, i, group=None, **kw):  # gloo rejects the 2-D output
        return orig_ag(o.view(-1), i.contiguous().view(-1), group=group, **kw)

    dist.all_gather_into_tensor = flat_ag
    from veomni.distributed.moe import moe_layer as ML

    T, E, K, H = 20 + 4 * rank, 8, 2, 16
    g = torch.Generator().manual_seed(200 + rank)
    hs = torch.randn(T, H, generator=g)
    logits = torch.randn(T, E, generator=g)
    rw, idx = torch.topk(torch.softmax(logits, -1), K, dim=-1)
    rw = rw / rw.sum(-1, keepdim=True)
    if rank == 0:
        idx = torch.where(idx == 3, torch.full_like(idx, 2), idx)  # rank 0 sends nothing to expert 3
    mask = torch.nn.functional.one_hot(idx, num_classes=E).permute(2, 1, 0)
    in_s, out_s, ngl, ngs = ML.preprocess(mask, E, group)
    tokens, routing_map, mapping, shape = ML.token_pre_all2all(hs, mask, E, in_s, out_s, ngl, group)
    # identity experts with a per-row marker so the combine can be checked exactly
    expert_out = tokens * 2.0
    final = ML.tokens_post_all2all(expert_out, rw, idx, E, in_s, out_s, ngl, routing_map, mapping, shape, group)
    dist.all_gather_into_tensor = orig_ag
    res["ep"] = {"hs": hs, "rw": rw, "idx": idx, "input_splits": in_s, "output_splits": out_s,
                 "num_global_tokens_per_local_expert": ngl.clone(), "num_global_sum_tokens_per_local_expert": ngs.clone(),
                 "tokens": tokens, "mapping": mapping, "final": final}

    # ---- FSDP2 on gloo: sharded grads after one backward -------------------------------------
    from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

    torch.manual_seed(7)
    model = torch.nn.Sequential(torch.nn.Linear(16, 24, bias=False), torch.nn.Linear(24, 8, bias=False))
    full = {n: p.detach().clone() for n, p in model.named_parameters()}
    mpp = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32)
    for layer in model:
        fully_shard(layer, mp_policy=mpp)
    fully_shard(model, mp_policy=mpp)
    g = torch.Generator().manual_seed(300 + rank)
    xin = torch.randn(5, 16, generator=g)
    model(xin).square().sum().backward()
    res["fsdp"] = {"full_params": full, "x": xin,
                   "sharded_grads": {n: p.grad.to_local().clone() for n, p in model.named_parameters()}}
    # ---- SP loss reduction (reference): token-weighted mean over the SP group + its backward ----
    from veomni.distributed.sequence_parallel import comm as SPC
    from veomni.distributed.sequence_parallel.loss import reduce_sequence_parallel_loss

    SPC.set_unified_sequence_parallel_group(group)
    cases = []
    for loss_v, n_valid in ((1.5 + rank, 7 + 3 * rank), (0.25, 0 if rank == 0 else 5), (2.0, 0)):
        loss_in = torch.tensor(float(loss_v), requires_grad=True)
        out = reduce_sequence_parallel_loss(loss_in * 1.0, torch.tensor(n_valid))
        (g_in,) = torch.autograd.grad(out * 3.0, loss_in)
        cases.append({"loss": float(loss_v), "n_valid": n_valid, "reduced": out.detach().clone(), "grad": g_in.clone()})
    res["sp_loss"] = cases
    torch.save(res, os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def gen_multirank():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, os.path.join(d, "store"), d), nprocs=world, join=True)
        ranks = [torch.load(os.path.join(d, f"rank{r}.pt"), weights_only=False) for r in range(world)]

    # Ulysses: oracle vs reference
    xs = [r["ulysses"]["x"] for r in ranks]
    gathered = o_comm.gather_seq_scatter_heads(xs, seq_dim=0, head_dim=1)
    for r in range(world):
        eq(gathered[r], ranks[r]["ulysses"]["gathered"], f"ulysses gather_seq_scatter_heads rank{r}")
    back = o_comm.gather_heads_scatter_seq(gathered, head_dim=1, seq_dim=0)
    for r in range(world):
        eq(back[r], xs[r], f"ulysses round trip rank{r}")

    # EP: oracle vs reference
    E = 8
    hs = [r["ep"]["hs"] for r in ranks]
    idx = [r["ep"]["idx"] for r in ranks]
    rw = [r["ep"]["rw"] for r in ranks]
    disp = o_moe.ep_dispatch(hs, idx, E)
    for r in range(world):
        ref = ranks[r]["ep"]
        assert disp[r]["input_splits"] == ref["input_splits"], "input_splits"
        assert disp[r]["output_splits"] == ref["output_splits"], "output_splits"
        eq(disp[r]["num_global_tokens_per_local_expert"], ref["num_global_tokens_per_local_expert"], f"ep counts rank{r}")
        eq(disp[r]["num_global_sum_tokens_per_local_expert"], ref["num_global_sum_tokens_per_local_expert"], f"ep sums rank{r}")
        eq(disp[r]["permutation_mapping"], ref["mapping"], f"ep permutation mapping rank{r}")
        eq(disp[r]["tokens"], ref["tokens"], f"ep dispatched tokens rank{r}")
    fin = o_moe.ep_combine([d["tokens"] * 2.0 for d in disp], disp, rw, idx, E, [h.shape for h in hs])
    for r in range(world):
        eq(fin[r], ranks[r]["ep"]["final"], f"ep combine rank{r}")

    # FSDP2: oracle reduce-scatter of bf16-compute grads == FSDP2's sharded grads
    full = ranks[0]["fsdp"]["full_params"]
    grads = []
    for r in range(world):
        w0 = full["0.weight"].to(torch.bfloat16).requires_grad_(True)
        w1 = full["1.weight"].to(torch.bfloat16).requires_grad_(True)
        x = ranks[r]["fsdp"]["x"].to(torch.bfloat16)
        torch.nn.functional.linear(torch.nn.functional.linear(x, w0), w1).square().sum().backward()
        grads.append({"0.weight": w0.grad, "1.weight": w1.grad})
    for name in ("0.weight", "1.weight"):
        shards = o_comm.fsdp_reduce_scatter([g[name] for g in grads], torch.float32, divide_factor=float(world))
        for r in range(world):
            ref = ranks[r]["fsdp"]["sharded_grads"][name]
            eq(shards[r].view(ref.shape), ref, f"fsdp2 reduce-scatter {name} rank{r}")
        ag = o_comm.fsdp_all_gather([c.reshape(-1) for c in full[name].chunk(world, dim=0)], torch.bfloat16)
        eq(ag.view(full[name].shape), full[name].to(torch.bfloat16), f"fsdp2 all-gather {name}")
    # SP loss reduction: oracle vs reference (forward value and the gradient of 3 * reduced w.r.t. each rank's loss)
    for c in range(len(ranks[0]["sp_loss"])):
        red, grads_sp = o_loss.reduce_sequence_parallel_loss([ranks[r]["sp_loss"][c]["loss"] for r in range(world)],
                                                             [ranks[r]["sp_loss"][c]["n_valid"] for r in range(world)], 3.0)
        for r in range(world):
            eq(torch.tensor(red), ranks[r]["sp_loss"][c]["reduced"], f"sp loss reduce case{c} rank{r}", atol=1e-6, rtol=1e-6)
            eq(torch.tensor(grads_sp[r]), ranks[r]["sp_loss"][c]["grad"], f"sp loss grad case{c} rank{r}", atol=1e-6, rtol=1e-6)

    for r in ranks:
        r["fsdp"]["bf16_grads"] = None
    for r in range(world):
        ranks[r]["fsdp"]["bf16_grads"] = grads[r]
    torch.save({"ranks": ranks}, HERE / "multirank.pt")


# ----------------------------------------------------------------------------------------------
def gen_qwen3_toy():
    """A toy Qwen3 causal LM step through the reference's own model builder (eager ops, fp32)."""
    sys.path.insert(0, "/root/reference/tests")
    from tools.training_utils import make_eager_ops_config  # reference test helper
    from transformers import Qwen3Config
    from veomni.models import build_foundation_model

    cfg = Qwen3Config(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=64, vocab_size=512, max_position_embeddings=512,
                      rms_norm_eps=1e-6, tie_word_embeddings=False, architectures=["Qwen3ForCausalLM"])
    with tempfile.TemporaryDirectory() as d:
        cfg.save_pretrained(d)
        torch.manual_seed(0)
        model = build_foundation_model(config_path=d, weights_path=None, torch_dtype="float32", init_device="cpu",
                                       ops_implementation=make_eager_ops_config())
    torch.manual_seed(0)
    model.init_weights()
    with torch.no_grad():  # non-trivial norm weights so q/k-norm parity means something
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    g = torch.Generator().manual_seed(5)
    lens = [40, 56, 32]
    ids = torch.randint(0, 512, (1, sum(lens)), generator=g)
    pos = torch.cat([torch.arange(n) for n in lens])[None]
    labels = ids.clone()
    off = 0
    for n in lens:  # DummyTextDataset convention: first label of each sample is ignored
        labels[0, off] = -100
        off += n
    # block-diagonal causal mask == what cu_seqlens tells flash-attn on the GPU path
    T = sum(lens)
    mask = torch.full((T, T), float("-inf"))
    off = 0
    for n in lens:
        mask[off : off + n, off : off + n] = torch.triu(torch.full((n, n), float("-inf")), diagonal=1)
        off += n
    out = model(input_ids=ids, position_ids=pos, attention_mask=mask[None, None], labels=labels, use_cache=False)
    out.loss.backward()
    gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in model.parameters()))
    print(f"  reference toy Qwen3: loss {out.loss.item():.6f} grad_norm {gn.item():.6f}")
    torch.save({
        "config": cfg.to_dict(), "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "input_ids": ids, "position_ids": pos, "labels": labels, "seq_lens": lens,
        "loss": out.loss.detach(), "grad_norm": gn.detach(),
        "grads": {k: p.grad.detach().clone() for k, p in model.named_parameters() if "layernorm" in k or "q_norm" in k or k.endswith("norm.weight")},
    }, HERE / "qwen3_toy.pt")


# ----------------------------------------------------------------------------------------------
def gen_loss():
    """ForCausalLMLoss bound to the reference's eager cross-entropy, on CPU: logits path and hidden+weights path."""
    from functools import partial

    from veomni.ops.kernels.cross_entropy import ForCausalLMLoss
    from veomni.ops.kernels.cross_entropy.eager import eager_cross_entropy

    loss_fn = partial(ForCausalLMLoss, cross_entropy_fn=eager_cross_entropy)
    g = torch.Generator().manual_seed(7)
    out = {}
    T, H, V = 70, 48, 1003  # odd vocabulary: exercises the scalar head/tail of the vectorised kernel
    labels = torch.randint(0, V, (1, T), generator=g)
    labels[0, 5:9] = -100
    labels[0, -3] = -100
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        for nib in (None, 50):
            key = f"{tag}/{'mean' if nib is None else 'sum'}"
            # (a) logits given
            logits = (torch.randn(1, T, V, generator=g) * 2.0).to(dtype).requires_grad_(True)
            loss, _, _ = loss_fn(logits=logits, labels=labels, vocab_size=V, num_items_in_batch=nib)
            (gl,) = torch.autograd.grad(loss, logits)
            sl = o_loss.shift_labels(labels).reshape(-1)
            lo = o_loss.cross_entropy(logits.detach().reshape(-1, V), sl, nib)
            eq(lo, loss.detach(), f"cross_entropy loss {key}", atol=1e-6, rtol=1e-6)
            scale = 1.0 / float((sl != -100).sum()) if nib is None else 1.0 / nib
            go = o_loss.cross_entropy_grad(logits.detach().reshape(-1, V), sl, scale).to(dtype)
            eq(go, gl.reshape(-1, V), f"cross_entropy grad {key}", atol=1e-7 if dtype == torch.float32 else 1e-4, rtol=1e-5 if dtype == torch.float32 else 8e-3)
            out[f"logits/{key}"] = {"logits": logits.detach(), "labels": labels, "num_items": nib, "loss": loss.detach(), "grad": gl}
            # (b) hidden states + lm_head weight (eager: F.linear(...).float())
            h = torch.randn(1, T, H, generator=g).to(dtype).requires_grad_(True)
            w = (torch.randn(V, H, generator=g) * 0.2).to(dtype).requires_grad_(True)
            loss, _, _ = loss_fn(labels=labels, vocab_size=V, num_items_in_batch=nib, hidden_states=h, weights=w)
            gh, gw = torch.autograd.grad(loss, (h, w))
            lo, dho, dwo = o_loss.fused_linear_cross_entropy(h.detach()[0], w.detach(), sl, nib, chunk_size=32)
            tol = dict(atol=1e-6, rtol=1e-5) if dtype == torch.float32 else dict(atol=2e-3, rtol=2e-2)
            eq(lo, loss.detach(), f"fused-linear loss {key}", atol=1e-5, rtol=1e-5)
            eq(dho, gh[0], f"fused-linear d hidden {key}", **tol)
            eq(dwo, gw, f"fused-linear d weight {key}", **tol)
            out[f"linear/{key}"] = {"hidden": h.detach(), "weight": w.detach(), "labels": labels, "num_items": nib,
                                    "loss": loss.detach(), "grad_hidden": gh, "grad_weight": gw}
    torch.save(out, HERE / "loss.pt")


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:] or ["ops", "moe", "multirank", "qwen3", "loss"]
    if "ops" in which:
        print("ops"); gen_ops()
    if "moe" in which:
        print("moe"); gen_moe_local()
    if "multirank" in which:
        print("multirank"); gen_multirank()
    if "qwen3" in which:
        print("qwen3 toy"); gen_qwen3_toy()
    if "loss" in which:
        print("loss"); gen_loss()
    print("golden fixtures written to", HERE)
