"""Additional microbenchmarks (attention, comm, MoE); imported lazily by tools/microbench.py."""
import json
import os

import torch

from tools.microbench import BF, report, time_fn


def bench_attention(dev, iters):
    from veomni_b200.attention import flash_attn_varlen

    T, Hq, Hk, D = 4096, 32, 8, 128
    sets = []
    for _ in range(3):
        sets.append(tuple(torch.randn(T, h, D, device=dev, dtype=BF) for h in (Hq, Hk, Hk)))
    cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
    flops_fwd = 4 * T * T * D * Hq / 2
    with torch.no_grad():
        report("attn_fwd[4096,32/8,128,causal]", time_fn(lambda q, k, v: flash_attn_varlen(q, k, v, cu, T), sets, iters),
               flops=flops_fwd)
    from veomni_b200 import attention as A

    if os.environ.get("VB200_PROFILE_ONCE", "0") == "1":  # ncu capture run: the default kernels only, once each
        gs = tuple(t.clone().requires_grad_(True) for t in sets[0])
        do = torch.randn(T, Hq, D, device=dev, dtype=BF)
        o = flash_attn_varlen(*gs, cu, T)
        report("attn_bwd only (delta + dQ + dK/dV)[4096,32/8,128,causal]",
               time_fn(lambda: o.backward(do, retain_graph=True), [()], iters), flops=flops_fwd * 2.5)
        return

    old = A.FWD_IMPL
    A.FWD_IMPL = "tc"
    try:
        with torch.no_grad():
            report("attn_fwd_tcgen05[4096,32/8,128,causal]", time_fn(lambda q, k, v: flash_attn_varlen(q, k, v, cu, T), sets, iters),
                   flops=flops_fwd)
    except Exception as ex:  # noqa: BLE001
        print({"attn_fwd_tcgen05": str(ex)})
    old_w8 = A.FWD_W8
    for w8 in (False, True):
        A.FWD_IMPL, A.FWD_W8 = "tc", w8
        try:
            with torch.no_grad():
                report(f"attn_fwd_tcgen05 (W8={w8})[4096,32/8,128,causal]",
                       time_fn(lambda q, k, v: flash_attn_varlen(q, k, v, cu, T), sets, iters), flops=flops_fwd)
        except Exception as ex:  # noqa: BLE001
            print({"attn_fwd_tcgen05_w8": str(ex)})
    A.FWD_W8 = old_w8
    A.FWD_IMPL = old
    gsets = [tuple(t.clone().requires_grad_(True) for t in s) for s in sets]
    do = torch.randn(T, Hq, D, device=dev, dtype=BF)

    def fb(q, k, v):
        o = flash_attn_varlen(q, k, v, cu, T)
        o.backward(do)
        q.grad = k.grad = v.grad = None

    report("attn_fwd+bwd[4096,32/8,128,causal]", time_fn(fb, gsets, iters), flops=flops_fwd * 3.5)
    old2 = (A.FWD_IMPL, A.BWD_IMPL)
    A.FWD_IMPL = A.BWD_IMPL = "tc"
    try:
        report("attn_fwd+bwd_tcgen05[4096,32/8,128,causal]", time_fn(fb, gsets, iters), flops=flops_fwd * 3.5)
    except Exception as ex:  # noqa: BLE001
        print({"attn_bwd_tcgen05": str(ex)})
    try:
        A.BWD_PP = not A.BWD_PP
        report(f"attn_fwd+bwd_tcgen05 (BWD_PP={A.BWD_PP})[4096,32/8,128,causal]", time_fn(fb, gsets, iters), flops=flops_fwd * 3.5)
    except Exception as ex:  # noqa: BLE001
        print({"attn_bwd_tcgen05_pp": str(ex)})
    A.BWD_PP = not A.BWD_PP
    for p16 in (False, True):
        A.BWD_P16 = p16
        try:
            report(f"attn_fwd+bwd_tcgen05 (BWD_P16={p16})[4096,32/8,128,causal]", time_fn(fb, gsets, iters), flops=flops_fwd * 3.5)
            qg, kg, vg = gsets[0]
            o = flash_attn_varlen(qg, kg, vg, cu, T)

            def bwd_only():
                o.backward(do, retain_graph=True)
                qg.grad = kg.grad = vg.grad = None

            report(f"attn_bwd only (delta + dQ + dK/dV) (BWD_P16={p16})[4096,32/8,128,causal]", time_fn(lambda: bwd_only(), [()], iters),
                   flops=flops_fwd * 2.5)
        except Exception as ex:  # noqa: BLE001
            print({"attn_bwd_p16": str(ex)})
    A.BWD_P16 = False
    for n128 in (False, True):
        A.BWD_DQ_N128 = n128
        try:
            qg, kg, vg = gsets[0]
            o = flash_attn_varlen(qg, kg, vg, cu, T)

            def bwd_only2():
                o.backward(do, retain_graph=True)
                qg.grad = kg.grad = vg.grad = None

            report(f"attn_bwd only (delta + dQ + dK/dV) (DQ_N128={n128})[4096,32/8,128,causal]", time_fn(lambda: bwd_only2(), [()], iters),
                   flops=flops_fwd * 2.5)
        except Exception as ex:  # noqa: BLE001
            print({"attn_bwd_dq_n128": str(ex)})
    A.BWD_DQ_N128 = True
    A.FWD_IMPL, A.BWD_IMPL = old2
    try:
        from flash_attn import flash_attn_varlen_func

        with torch.no_grad():
            report("(lib) flash_attn2_fwd", time_fn(lambda q, k, v: flash_attn_varlen_func(q, k, v, cu, cu, T, T, causal=True), sets, iters), flops=flops_fwd)

        def fb2(q, k, v):
            o = flash_attn_varlen_func(q, k, v, cu, cu, T, T, causal=True)
            o.backward(do)
            q.grad = k.grad = v.grad = None

        report("(lib) flash_attn2_fwd+bwd", time_fn(fb2, gsets, iters), flops=flops_fwd * 3.5)
    except Exception as ex:  # noqa: BLE001
        print({"flash_attn": str(ex)})


def bench_moe_ep8(dev, iters):
    """GroupGEMM at the per-rank shape of BASELINE configs[3] (Qwen3-30B-A3B, EP8): 16 local experts, the T*K = 32768 rows a
    rank receives on average (multinomial split, ~2048 rows per expert), H=2048, I=768. Unlike the 128-local-expert shape
    of ``bench_moe`` (805 MB of weights per 206 GFLOP: below the HBM ridge) this one is tensor-bound."""
    from veomni_b200.moe import group_gemm_same_mn, group_gemm_same_nk

    E, R, H, I = 16, 32768, 2048, 768
    g = torch.Generator(device=dev).manual_seed(1)
    counts = torch.bincount(torch.randint(0, E, (R,), device=dev, generator=g), minlength=E)
    cumsum = torch.cumsum(counts, 0).to(torch.int32)
    x = (0.1 * torch.randn(R, H, device=dev, generator=g)).to(BF)
    w1 = (0.1 * torch.randn(E, 2 * I, H, device=dev, generator=g)).to(BF)
    w2 = (0.1 * torch.randn(E, H, I, device=dev, generator=g)).to(BF)
    with torch.no_grad():
        report("ep8 group_gemm_NT fc1[32768x2048 -> 1536, 16e]", time_fn(lambda: group_gemm_same_nk(x, w1, cumsum, transpose_b=True), [()], iters),
               flops=2 * R * 2 * I * H)
        a = group_gemm_same_nk(x, w1, cumsum, transpose_b=True)
        act = a[:, :I].contiguous()
        report("ep8 group_gemm_NT fc2[32768x768 -> 2048, 16e]", time_fn(lambda: group_gemm_same_nk(act, w2, cumsum, transpose_b=True), [()], iters),
               flops=2 * R * I * H)
        report("ep8 group_gemm_NN dgrad fc1[32768x1536 -> 2048, 16e]", time_fn(lambda: group_gemm_same_nk(a, w1, cumsum, transpose_b=False), [()], iters),
               flops=2 * R * 2 * I * H)
        gw = torch.empty_like(w1)
        report("ep8 group_gemm_TN wgrad fc1[16 x 1536x2048, K=32768]", time_fn(lambda: group_gemm_same_mn(a, x, gw, cumsum), [()], iters),
               flops=2 * R * 2 * I * H)
        gw2 = torch.empty_like(w2)
        go = (0.1 * torch.randn(R, H, device=dev, generator=g)).to(BF)
        report("ep8 group_gemm_TN wgrad fc2[16 x 2048x768, K=32768]", time_fn(lambda: group_gemm_same_mn(go, act, gw2, cumsum), [()], iters),
               flops=2 * R * I * H)


BENCHES = {"attention": bench_attention, "moe_ep8": bench_moe_ep8}


def bench_moe(dev, iters):
    """Qwen3-30B-A3B layer shapes: T=4096 tokens, top-8 of 128 experts, H=2048, I=768 (T*K = 32768 rows)."""
    from veomni_b200.moe import fused_moe_forward, group_gemm_same_mn, group_gemm_same_nk, moe_gather, moe_route, moe_scatter

    T, E, K, H, I = 4096, 128, 8, 2048, 768
    g = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(T, E, device=dev, generator=g)
    rw, idx = torch.topk(torch.softmax(logits, -1), K, dim=-1)
    rw = (rw / rw.sum(-1, keepdim=True)).to(BF)
    hs = (0.1 * torch.randn(T, H, device=dev, generator=g)).to(BF)
    w1 = (0.1 * torch.randn(E, 2 * I, H, device=dev, generator=g)).to(BF)
    w2 = (0.1 * torch.randn(E, H, I, device=dev, generator=g)).to(BF)
    with torch.no_grad():
        report("moe_route[32768 slots,128e]", time_fn(lambda: moe_route(idx, E), [()], iters))
        splits, cumsum, sidx = moe_route(idx, E)
        report("moe_scatter[4096x8x2048]", time_fn(lambda: moe_scatter(hs, sidx), [()], iters), nbytes=(T + T * K) * H * 2)
        x = moe_scatter(hs, sidx)
        report("moe_gather[4096x8x2048]", time_fn(lambda: moe_gather(x, sidx), [()], iters), nbytes=(T + T * K) * H * 2)
        report("group_gemm_NT fc1[32768x2048 -> 1536]", time_fn(lambda: group_gemm_same_nk(x, w1, cumsum, transpose_b=True), [()], iters),
               flops=2 * T * K * 2 * I * H)
        a = group_gemm_same_nk(x, w1, cumsum, transpose_b=True)
        act = a[:, :I].contiguous()
        report("group_gemm_NT fc2[32768x768 -> 2048]", time_fn(lambda: group_gemm_same_nk(act, w2, cumsum, transpose_b=True), [()], iters),
               flops=2 * T * K * I * H)
        report("group_gemm_NN dgrad fc1[32768x1536 -> 2048]", time_fn(lambda: group_gemm_same_nk(a, w1, cumsum, transpose_b=False), [()], iters),
               flops=2 * T * K * 2 * I * H)
        gw = torch.empty_like(w1)
        report("group_gemm_TN wgrad fc1[128 x 1536x2048, K=32768]", time_fn(lambda: group_gemm_same_mn(a, x, gw, cumsum), [()], iters),
               flops=2 * T * K * 2 * I * H)
        gw2 = torch.empty_like(w2)
        go = (0.1 * torch.randn(T * K, H, device=dev, generator=g)).to(BF)
        report("group_gemm_TN wgrad fc2[128 x 2048x768, K=32768]", time_fn(lambda: group_gemm_same_mn(go, act, gw2, cumsum), [()], iters),
               flops=2 * T * K * I * H)
    hs_g, w1_g, w2_g, rw_g = (t.clone().requires_grad_(True) for t in (hs, w1, w2, rw))

    def fb():
        out = fused_moe_forward(E, rw_g, idx, hs_g, None, None, w2_g, fc1_1_2_weight=w1_g)
        out.backward(hs)
        hs_g.grad = w1_g.grad = w2_g.grad = rw_g.grad = None

    report("fused_moe fwd+bwd[Qwen3-30B-A3B layer, T=4096]", time_fn(fb, [()], iters), flops=3 * 2 * T * K * 3 * I * H)
    if os.environ.get("VB200_SKIP_QUACK", "0") == "1":
        return
    try:  # the existing Blackwell kernel the reference can call (veomni/ops/kernels/moe/quack_gemm.py:28,88-161): quack-kernels
        from quack.gemm_interface import gemm as qgemm

        cu = torch.zeros(E + 1, dtype=torch.int32, device=dev)
        cu[1:] = cumsum.int()
        w1t, w2t = w1.transpose(1, 2), w2.transpose(1, 2)
        with torch.no_grad():
            ref = qgemm(x, w1t, cu_seqlens_m=cu)
            mine = group_gemm_same_nk(x, w1, cumsum, transpose_b=True)
            print(json.dumps({"quack_vs_vb200_fc1_maxabs": float((ref.float() - mine.float()).abs().max()),
                              "ref_absmax": float(ref.float().abs().max())}), flush=True)
            report("(lib) quack gemm fc1 cu_seqlens_m[32768x2048 -> 1536]", time_fn(lambda: qgemm(x, w1t, cu_seqlens_m=cu), [()], iters),
                   flops=2 * T * K * 2 * I * H)
            report("(lib) quack gemm fc2 cu_seqlens_m[32768x768 -> 2048]", time_fn(lambda: qgemm(act, w2t, cu_seqlens_m=cu), [()], iters),
                   flops=2 * T * K * I * H)
            report("(lib) quack gemm dgrad fc1 cu_seqlens_m[32768x1536 -> 2048]", time_fn(lambda: qgemm(a, w1, cu_seqlens_m=cu), [()], iters),
                   flops=2 * T * K * 2 * I * H)
            report("(lib) quack gemm wgrad fc1 cu_seqlens_k[128 x 1536x2048]", time_fn(lambda: qgemm(a.T, x, cu_seqlens_k=cu), [()], iters),
                   flops=2 * T * K * 2 * I * H)
    except Exception as ex:  # noqa: BLE001
        print(json.dumps({"quack": f"{type(ex).__name__}: {str(ex)[:300]}"}), flush=True)
    try:  # library reference: cuBLAS dense GEMM of the same FLOPs
        a2 = torch.randn(T * K, H, device=dev, dtype=BF)
        b2 = torch.randn(2 * I, H, device=dev, dtype=BF)
        with torch.no_grad():
            report("(lib) cublas dense [32768x2048]x[2048x1536]", time_fn(lambda: torch.nn.functional.linear(a2, b2), [()], iters),
                   flops=2 * T * K * 2 * I * H)
    except Exception as ex:  # noqa: BLE001
        print({"cublas": str(ex)})


BENCHES["moe"] = bench_moe
