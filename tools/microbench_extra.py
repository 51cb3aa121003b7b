"""Additional microbenchmarks (attention, comm, MoE); imported lazily by tools/microbench.py."""
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

    old = A.FWD_IMPL
    A.FWD_IMPL = "tc"
    try:
        with torch.no_grad():
            report("attn_fwd_tcgen05[4096,32/8,128,causal]", time_fn(lambda q, k, v: flash_attn_varlen(q, k, v, cu, T), sets, iters),
                   flops=flops_fwd)
    except Exception as ex:  # noqa: BLE001
        print({"attn_fwd_tcgen05": str(ex)})
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


BENCHES = {"attention": bench_attention}
