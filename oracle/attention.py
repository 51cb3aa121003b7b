"""Oracle for packed (varlen) causal attention.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.

The reference calls flash-attn's ``flash_attn_varlen_func`` through HF's
``_flash_attention_forward`` (veomni/ops/kernels/attention/__init__.py:304-320; HF
transformers/modeling_flash_attention_utils.py padding-free branch).  flash-attn is third-party
CUDA (flash-attn 2.8.3, uv.lock:742-758) and cannot run on CPU, so the algorithm is restated from
its definition — per-sequence causal softmax(QK^T * scale) V with GQA head replication — exactly
as the reference's own eager path does (eager_attention_forward + repeat_kv,
veomni/models/transformers/qwen3/generated/patched_modeling_qwen3_gpu.py:226-261).  The reference
tests that pin this boundary are tests/models/test_models_patch.py (eager vs flash-attn, 1e-2) and
tests/parallel/ulysses/test_ulysses.py (SDPA inside the SP wrapper).
"""

from __future__ import annotations

import math

import torch


def varlen_causal_attention(q, k, v, cu_seqlens, scale: float | None = None, causal: bool = True):
    """q: [T, Hq, D], k/v: [T, Hkv, D] (packed), cu_seqlens: int tensor [nseq+1].

    Returns (out [T, Hq, D] in q.dtype, lse [Hq, T] fp32) — the log-sum-exp of the scaled scores is
    what flash-attn returns as ``softmax_lse`` and what the backward kernel consumes.
    Math in fp32 on the given (possibly bf16-rounded) inputs; output rounded once.
    """
    T, Hq, D = q.shape
    Hk = k.shape[1]
    rep = Hq // Hk
    scale = 1.0 / math.sqrt(D) if scale is None else scale
    out = torch.zeros(T, Hq, D, dtype=torch.float32)
    lse = torch.zeros(Hq, T, dtype=torch.float32)
    cu = [int(c) for c in cu_seqlens.tolist()]
    for a, b in zip(cu[:-1], cu[1:]):
        if b <= a:
            continue
        qs = q[a:b].float().transpose(0, 1)  # [Hq, L, D]
        ks = k[a:b].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        vs = v[a:b].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        s = torch.matmul(qs, ks.transpose(1, 2)) * scale
        if causal:
            L = b - a
            mask = torch.ones(L, L, dtype=torch.bool).tril()
            s = s.masked_fill(~mask, float("-inf"))
        l = torch.logsumexp(s, dim=-1)
        p = torch.exp(s - l[..., None])
        out[a:b] = torch.matmul(p, vs).transpose(0, 1)
        lse[:, a:b] = l
    return out.to(q.dtype), lse


def varlen_causal_attention_bwd(q, k, v, cu_seqlens, dout, scale: float | None = None, causal: bool = True):
    """Gradients of :func:`varlen_causal_attention` w.r.t. q, k, v via autograd in fp32."""
    qf = q.detach().float().requires_grad_(True)
    kf = k.detach().float().requires_grad_(True)
    vf = v.detach().float().requires_grad_(True)
    T, Hq, D = q.shape
    Hk = k.shape[1]
    rep = Hq // Hk
    scale = 1.0 / math.sqrt(D) if scale is None else scale
    outs = []
    cu = [int(c) for c in cu_seqlens.tolist()]
    for a, b in zip(cu[:-1], cu[1:]):
        if b <= a:
            continue
        qs = qf[a:b].transpose(0, 1)
        ks = kf[a:b].transpose(0, 1).repeat_interleave(rep, dim=0)
        vs = vf[a:b].transpose(0, 1).repeat_interleave(rep, dim=0)
        s = torch.matmul(qs, ks.transpose(1, 2)) * scale
        if causal:
            L = b - a
            s = s.masked_fill(~torch.ones(L, L, dtype=torch.bool).tril(), float("-inf"))
        outs.append(torch.matmul(torch.softmax(s, dim=-1), vs).transpose(0, 1))
    o = torch.cat(outs, dim=0)
    o.backward(dout.float()[: o.shape[0]])
    return qf.grad.to(q.dtype), kf.grad.to(k.dtype), vf.grad.to(v.dtype)
