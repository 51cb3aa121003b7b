"""The CPU oracle reproduces the reference outputs stored in tests/golden/*.pt.

The fixtures were produced by tests/golden/make_golden.py from the reference's own functions
(imported from /root/reference in the authoring container). Bit-exact unless stated.
"""
import torch

from oracle import attention as o_attn
from oracle import comm as o_comm
from oracle import moe as o_moe
from oracle import ops as o_ops


def test_rms_norm_matches_reference(golden):
    g = golden("ops.pt")
    for key in [k for k in g if k.startswith("rms_norm/")]:
        f = g[key]
        assert torch.equal(o_ops.rms_norm(f["x"], f["w"], f["eps"]), f["y"]), key


def test_rope_matches_reference(golden):
    g = golden("ops.pt")
    for tag, dtype in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        f = g[f"rope/{tag}"]
        cos, sin = o_ops.rotary_cos_sin(f["pos"], f["q"].shape[-1], f["theta"], dtype)
        assert torch.equal(cos, f["cos"]) and torch.equal(sin, f["sin"])
        q, k = o_ops.apply_rotary_pos_emb(f["q"], f["k"], f["cos"], f["sin"])
        assert torch.equal(q, f["q_out"]) and torch.equal(k, f["k_out"])


def test_swiglu_matches_reference(golden):
    g = golden("ops.pt")
    for tag in ("bf16", "fp32"):
        f = g[f"swiglu/{tag}"]
        assert torch.equal(o_ops.silu_mul(f["gate"], f["up"]), f["act"])
        assert torch.equal(o_ops.swiglu_mlp(f["x"], f["wg"], f["wu"], f["wd"]), f["y"])


def test_varlen_attention_matches_reference_eager(golden):
    f = golden("ops.pt")["attention/fp32"]
    out, lse = o_attn.varlen_causal_attention(f["q"], f["k"], f["v"], f["cu"])
    torch.testing.assert_close(out, f["out"], atol=2e-6, rtol=1e-5)
    torch.testing.assert_close(lse, f["lse"], atol=1e-6, rtol=1e-6)


def test_moe_matches_reference(golden):
    g = golden("moe.pt")
    f = g["moe/fp32"]
    E = f["gate_up"].shape[0]
    assert torch.equal(o_moe.eager_moe_forward(E, f["rw"], f["idx"], f["hs"], f["gate_up"], f["down"]), f["y"])
    y, inter = o_moe.fused_moe_forward(E, f["rw"], f["idx"], f["hs"], f["gate_up"], f["down"])
    torch.testing.assert_close(y, f["y"], atol=1e-5, rtol=1e-4)
    assert torch.equal(inter["scatter_index"], f["scatter_index"])
    assert torch.equal(inter["splits"], f["splits"])
    assert int(inter["splits"][5]) == 0  # the empty-expert edge case is in the fixture
    u = g["moe_utils/fp32"]
    p, m = o_moe.permute(f["hs"], u["routing_map"])
    assert torch.equal(p, u["perm"]) and torch.equal(m, u["mapping"])
    w = o_moe.generate_weights_idx(f["rw"], f["idx"], E)
    assert torch.equal(w, u["weights_idx"])
    assert torch.equal(o_moe.unpermute(p, w, f["hs"].shape, m, u["routing_map"]), u["unperm"])


def test_scatter_index_is_a_stable_rank():
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, 16, (257, 4), generator=g)
    s = o_moe.scatter_index(idx).flatten().long()
    assert sorted(s.tolist()) == list(range(idx.numel()))
    flat = idx.flatten()
    inv = torch.empty_like(s)
    inv[s] = torch.arange(s.numel())
    sorted_experts = flat[inv]
    assert torch.all(sorted_experts[1:] >= sorted_experts[:-1])  # expert-sorted
    same = sorted_experts[1:] == sorted_experts[:-1]
    assert torch.all(inv[1:][same] > inv[:-1][same])  # stable within an expert


def test_ulysses_matches_reference(golden):
    ranks = golden("multirank.pt")["ranks"]
    xs = [r["ulysses"]["x"] for r in ranks]
    got = o_comm.gather_seq_scatter_heads(xs, seq_dim=0, head_dim=1)
    for r, rk in enumerate(ranks):
        assert torch.equal(got[r], rk["ulysses"]["gathered"])
    back = o_comm.gather_heads_scatter_seq(got, head_dim=1, seq_dim=0)
    for r in range(len(ranks)):
        assert torch.equal(back[r], xs[r])


def test_ulysses_pads_ragged_sequence():
    xs = [torch.randn(7, 4, 8), torch.randn(7, 4, 8)]  # heads-sharded [S=7, H/P... ] with odd S
    out = o_comm.gather_heads_scatter_seq(xs, head_dim=1, seq_dim=0)
    assert out[0].shape == (4, 8, 8) and out[1].shape == (4, 8, 8)
    assert torch.count_nonzero(out[1][3]) == 0  # zero padding lands at the end of the last rank


def test_ep_dispatch_combine_matches_reference(golden):
    ranks = golden("multirank.pt")["ranks"]
    E = 8
    hs = [r["ep"]["hs"] for r in ranks]
    idx = [r["ep"]["idx"] for r in ranks]
    rw = [r["ep"]["rw"] for r in ranks]
    disp = o_moe.ep_dispatch(hs, idx, E)
    for r, rk in enumerate(ranks):
        ref = rk["ep"]
        assert disp[r]["input_splits"] == ref["input_splits"]
        assert disp[r]["output_splits"] == ref["output_splits"]
        assert torch.equal(disp[r]["num_global_tokens_per_local_expert"], ref["num_global_tokens_per_local_expert"])
        assert torch.equal(disp[r]["permutation_mapping"], ref["mapping"])
        assert torch.equal(disp[r]["tokens"], ref["tokens"])
    fin = o_moe.ep_combine([d["tokens"] * 2.0 for d in disp], disp, rw, idx, E, [h.shape for h in hs])
    for r, rk in enumerate(ranks):
        assert torch.equal(fin[r], rk["ep"]["final"])


def test_fsdp_collectives_match_reference(golden):
    ranks = golden("multirank.pt")["ranks"]
    world = len(ranks)
    full = ranks[0]["fsdp"]["full_params"]
    for name in full:
        shards = o_comm.fsdp_reduce_scatter([r["fsdp"]["bf16_grads"][name] for r in ranks], torch.float32, float(world))
        for r, rk in enumerate(ranks):
            ref = rk["fsdp"]["sharded_grads"][name]
            assert torch.equal(shards[r].view(ref.shape), ref)
        ag = o_comm.fsdp_all_gather([c.reshape(-1) for c in full[name].chunk(world, dim=0)], torch.bfloat16)
        assert torch.equal(ag.view(full[name].shape), full[name].to(torch.bfloat16))


def test_causal_lm_loss_matches_reference(golden):
    """ForCausalLMLoss + eager_cross_entropy (logits path) and the chunked fused-linear restatement."""
    from oracle import loss as o_loss

    g = golden("loss.pt")
    for key, f in g.items():
        kind, tag, red = key.split("/")
        sl = o_loss.shift_labels(f["labels"]).reshape(-1)
        fp32 = tag == "fp32"
        if kind == "logits":
            V = f["logits"].shape[-1]
            x = f["logits"].reshape(-1, V)
            loss = o_loss.cross_entropy(x, sl, f["num_items"])
            torch.testing.assert_close(loss, f["loss"], atol=1e-6, rtol=1e-6, msg=key)
            scale = 1.0 / float((sl != -100).sum()) if f["num_items"] is None else 1.0 / f["num_items"]
            grad = o_loss.cross_entropy_grad(x, sl, scale).to(x.dtype)
            torch.testing.assert_close(grad, f["grad"].reshape(-1, V), atol=1e-7 if fp32 else 1e-4, rtol=1e-5 if fp32 else 8e-3, msg=key)
            assert torch.all(grad[sl == -100] == 0)
        else:
            loss, dh, dw = o_loss.fused_linear_cross_entropy(f["hidden"][0], f["weight"], sl, f["num_items"], chunk_size=32)
            tol = dict(atol=1e-6, rtol=1e-5) if fp32 else dict(atol=2e-3, rtol=2e-2)
            torch.testing.assert_close(loss, f["loss"], atol=1e-5, rtol=1e-5, msg=key)
            torch.testing.assert_close(dh, f["grad_hidden"][0], msg=key, **tol)
            torch.testing.assert_close(dw, f["grad_weight"], msg=key, **tol)
