"""Packed (varlen) causal attention on the sm_100a kernels, and the SP-aware wrapper that mirrors
VeOmni's ``flash_attention_forward`` (veomni/ops/kernels/attention/__init__.py:151-332).
"""

from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from ._lib import VB200Error, check, stream_ptr


# When set to a list, every forward launch appends a (start, end) CUDA-event pair recorded on the
# launching stream (bench.py uses it for the live per-launch duration of the dominant kernel).
PROFILE = None


def _strides(*tensors):
    vals = []
    for t in tensors:
        if t.stride(-1) != 1:
            raise VB200Error("attention tensors must be contiguous in head_dim")
        vals += [t.stride(0), t.stride(1)]
    return (ctypes.c_int64 * len(vals))(*vals)


def _prep(t: torch.Tensor) -> torch.Tensor:
    """[T, H, D] tensor usable by TMA: 16-byte aligned base and strides."""
    if t.stride(-1) != 1 or t.data_ptr() % 16 or t.stride(0) % 8 or t.stride(1) % 8:
        t = t.contiguous()
    return t


class _VarlenAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, scale, causal):
        for t in (q, k, v):
            if not t.is_cuda or t.dtype != torch.bfloat16:
                raise VB200Error("veomni_b200 attention expects CUDA bfloat16 q/k/v (no CPU fallback)")
        q, k, v = _prep(q), _prep(k), _prep(v)
        T, Hq, D = q.shape
        Hk = k.shape[1]
        cu = cu_seqlens.to(device=q.device, dtype=torch.int32).contiguous()
        nseq = cu.numel() - 1
        o = torch.empty(T, Hq, D, dtype=q.dtype, device=q.device)
        lse = torch.empty(Hq, T, dtype=torch.float32, device=q.device)
        scale = float(scale) if scale is not None else 1.0 / math.sqrt(D)
        lib = _lib.load()
        with torch.cuda.device(q.device):
            if PROFILE is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            check(
                lib.vb200_attn_varlen_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                          cu.data_ptr(), nseq, int(max_seqlen), T, Hq, Hk, D, _strides(q, k, v, o),
                                          scale, 1 if causal else 0, stream_ptr()),
                "vb200_attn_varlen_fwd",
            )
            if PROFILE is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                PROFILE.append((ev0, ev1))
        ctx.save_for_backward(q, k, v, o, lse, cu)
        ctx.meta = (int(max_seqlen), scale, bool(causal))
        return o

    @staticmethod
    def backward(ctx, dout):
        q, k, v, o, lse, cu = ctx.saved_tensors
        max_seqlen, scale, causal = ctx.meta
        dout = _prep(dout)
        T, Hq, D = q.shape
        Hk = k.shape[1]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        if dq.stride() != q.contiguous().stride():
            dq, dk, dv = (torch.empty(t.shape, dtype=t.dtype, device=t.device) for t in (q, k, v))
        delta = torch.empty(Hq, T, dtype=torch.float32, device=q.device)
        lib = _lib.load()
        with torch.cuda.device(q.device):
            check(
                lib.vb200_attn_varlen_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(),
                                          lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                          cu.data_ptr(), cu.numel() - 1, max_seqlen, T, Hq, Hk, D,
                                          _strides(q, k, v, o, dout, dq, dk, dv), scale, 1 if causal else 0,
                                          stream_ptr()),
                "vb200_attn_varlen_bwd",
            )
        return dq, dk, dv, None, None, None, None


def flash_attn_varlen(q, k, v, cu_seqlens, max_seqlen: int, softmax_scale: float | None = None, causal: bool = True):
    """q ``[T,Hq,D]``, k/v ``[T,Hkv,D]`` packed bf16; ``cu_seqlens`` int32 ``[nseq+1]``. Returns ``[T,Hq,D]``."""
    return _VarlenAttn.apply(q, k, v, cu_seqlens, max_seqlen, softmax_scale, causal)
