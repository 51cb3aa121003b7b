"""Packed (varlen) causal attention on the sm_100a kernels, and the SP-aware wrapper that mirrors
VeOmni's ``flash_attention_forward`` (veomni/ops/kernels/attention/__init__.py:151-332).
"""

from __future__ import annotations

import ctypes
import math
import os

import torch

from . import _lib
from ._lib import VB200Error, check, stream_ptr


# When set to a dict {"fwd": [], "bwd_dq": [], "bwd_dkdv": []}, every launch of the corresponding kernel appends a
# (start, end) CUDA-event pair recorded on the launching stream (bench.py: live per-launch durations).  A list is
# accepted for the forward kernel only.
PROFILE = None


def _prof(tag):
    if PROFILE is None:
        return None
    lst = PROFILE if isinstance(PROFILE, list) and tag == "fwd" else (PROFILE.get(tag) if isinstance(PROFILE, dict) else None)
    if lst is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return lst, ev


def _prof_end(h):
    if h is not None:
        lst, ev0 = h
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        lst.append((ev0, ev1))

# Forward implementation for head_dim 128: "tc" = tcgen05/TMEM pipeline (attention_tc.cu), "mma" = mma.sync kernel.
FWD_IMPL = os.environ.get("VB200_ATTN_FWD", "tc")
BWD_IMPL = os.environ.get("VB200_ATTN_BWD", "tc")
FWD_W8 = os.environ.get("VB200_ATTN_FWD_W8", "1") == "1"  # eight softmax warps (validated on B200: 173 vs 178 us at T=4096)
BWD_PP = os.environ.get("VB200_ATTN_BWD_PP", "0") == "1"  # softmax warpgroups on alternate tiles ("ping-pong")
BWD_DQ_N128 = os.environ.get("VB200_ATTN_BWD_DQ_N128", "1") == "1"  # dQ kernel with 128-row K/V tiles (N = 128 MMAs); bit-identical, -25 % (B200)
BWD_TRACE = False  # debugging: record the dQ kernel's hand-off timeline (tools/attn_trace.py)
BWD_P16 = os.environ.get("VB200_ATTN_BWD_P16", "0") == "1"  # sixteen softmax warps: two groups of eight on alternate tiles
BWD_DQ_SS = False  # True: dQ kernel with every MMA operand in shared memory (cross-check of the A-in-TMEM default)


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
    def forward(ctx, q, k, v, cu_seqlens, max_seqlen, scale, causal, replay=None):
        for t in (q, k, v):
            if not t.is_cuda or t.dtype != torch.bfloat16:
                raise VB200Error("veomni_b200 attention expects CUDA bfloat16 q/k/v (no CPU fallback)")
        q, k, v = _prep(q), _prep(k), _prep(v)
        T, Hq, D = q.shape
        Hk = k.shape[1]
        cu = cu_seqlens.to(device=q.device, dtype=torch.int32).contiguous()
        nseq = cu.numel() - 1
        scale = float(scale) if scale is not None else 1.0 / math.sqrt(D)
        if replay is not None:
            # gradient-checkpoint recompute: the kernel is deterministic, so the (o, lse) kept from the first
            # forward ARE what a relaunch would produce — keep them resident instead of recomputing.
            o, lse = replay
            ctx.save_for_backward(q, k, v, o, lse, cu)
            ctx.meta = (int(max_seqlen), scale, bool(causal))
            return o
        o = torch.empty(T, Hq, D, dtype=q.dtype, device=q.device)
        lse = torch.empty(Hq, T, dtype=torch.float32, device=q.device)
        lib = _lib.load()
        with torch.cuda.device(q.device):
            hprof = _prof("fwd")
            use_tc = FWD_IMPL == "tc" and D == 128
            fwd = lib.vb200_attn_varlen_fwd_tc if use_tc else lib.vb200_attn_varlen_fwd
            flags = (1 if causal else 0) | ((1 << 8) if (use_tc and FWD_W8) else 0)
            check(
                fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), cu.data_ptr(), nseq,
                    int(max_seqlen), T, Hq, Hk, D, _strides(q, k, v, o), scale, flags, stream_ptr()),
                "vb200_attn_varlen_fwd",
            )
            _prof_end(hprof)
        ctx.save_for_backward(q, k, v, o, lse, cu)
        ctx.meta = (int(max_seqlen), scale, bool(causal))
        ctx.mark_non_differentiable(lse)
        return o, lse

    @staticmethod
    def backward(ctx, dout, _dlse=None):
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
        if BWD_IMPL == "tc" and D == 128:
            with torch.cuda.device(q.device):
                check(lib.vb200_attn_bwd_delta(o.data_ptr(), dout.data_ptr(), delta.data_ptr(), T, Hq, D, o.stride(0), o.stride(1),
                                               dout.stride(0), dout.stride(1), stream_ptr()), "vb200_attn_bwd_delta")
                strides = _strides(q, k, v, dout, dq, dk, dv)
                for tag, only in (("bwd_dq", 1), ("bwd_dkdv", 2)) if isinstance(PROFILE, dict) else ((None, 0),):
                    hprof = _prof(tag) if tag else None
                    check(lib.vb200_attn_varlen_bwd_tc(q.data_ptr(), k.data_ptr(), v.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                                       delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                                       cu.data_ptr(), cu.numel() - 1, max_seqlen, T, Hq, Hk, D, strides, scale,
                                                       (1 if causal else 0) | (only << 8) | ((1 << 10) if BWD_DQ_SS else 0) | ((1 << 11) if BWD_PP else 0) | ((1 << 12) if BWD_P16 else 0) | ((1 << 13) if BWD_TRACE else 0) | ((1 << 14) if BWD_DQ_N128 else 0),
                                                       stream_ptr()),
                          "vb200_attn_varlen_bwd_tc")
                    _prof_end(hprof)
            return dq, dk, dv, None, None, None, None, None
        with torch.cuda.device(q.device):
            check(
                lib.vb200_attn_varlen_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(),
                                          lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                          cu.data_ptr(), cu.numel() - 1, max_seqlen, T, Hq, Hk, D,
                                          _strides(q, k, v, o, dout, dq, dk, dv), scale, 1 if causal else 0,
                                          stream_ptr()),
                "vb200_attn_varlen_bwd",
            )
        return dq, dk, dv, None, None, None, None, None


def flash_attn_varlen(q, k, v, cu_seqlens, max_seqlen: int, softmax_scale: float | None = None, causal: bool = True,
                      return_lse: bool = False, replay=None):
    """q ``[T,Hq,D]``, k/v ``[T,Hkv,D]`` packed bf16; ``cu_seqlens`` int32 ``[nseq+1]``. Returns ``[T,Hq,D]``
    (and the fp32 log-sum-exp ``[Hq,T]`` with ``return_lse``).  ``replay=(o, lse)`` re-attaches a previously computed
    result to the autograd graph without relaunching the forward kernel (checkpoint recompute)."""
    out = _VarlenAttn.apply(q, k, v, cu_seqlens, max_seqlen, softmax_scale, causal, replay)
    if replay is not None:
        return (out, replay[1]) if return_lse else out
    return out if return_lse else out[0]


def flash_attention_forward(module, query, key, value, attention_mask, dropout: float = 0.0, scaling: float | None = None,
                            sliding_window=None, softcap=None, skip_ulysses: bool = False, **kwargs):
    """Drop-in for VeOmni's ``flash_attention_forward`` (veomni/ops/kernels/attention/__init__.py:151-332), the
    function registered in HF ``ALL_ATTENTION_FUNCTIONS`` for the ``veomni_flash_attention_*_with_sp`` names.

    query ``[B,Hq,S,D]``, key/value ``[B,Hkv,S,D]`` (HF layout) -> ``(attn_output [B,S,Hq,D], None)``.
    Packed ("padding-free") batches only — VeOmni's collator always supplies ``cu_seq_lens_q/k`` and
    ``max_length_q/k`` (veomni/data/data_collator.py:44-73) — with the Ulysses gather/scatter around the
    kernel exactly where the reference has it (:232-281, :322-330).
    """
    if sliding_window is not None or softcap is not None or dropout:
        raise VB200Error("veomni_b200 attention: sliding_window / softcap / dropout are not supported")
    cu = kwargs.get("cu_seq_lens_q")
    if cu is None or query.shape[0] != 1:
        raise VB200Error("veomni_b200 attention needs a packed batch (B == 1) with cu_seq_lens_q/max_length_q kwargs")
    max_len = int(kwargs.get("max_length_q") or 0)
    is_causal = kwargs.pop("is_causal", None)  # reference :228-230: an explicit None falls back to module.is_causal
    if is_causal is None:
        is_causal = getattr(module, "is_causal", True)
    q = query.transpose(1, 2).squeeze(0)  # [S, Hq, D]
    k = key.transpose(1, 2).squeeze(0)
    v = value.transpose(1, 2).squeeze(0)
    group = None
    if not skip_ulysses:
        try:
            from veomni.distributed.parallel_state import get_parallel_state

            ps = get_parallel_state()
            if ps.ulysses_enabled:
                group = ps.ulysses_group
        except ImportError:
            group = getattr(module, "ulysses_group", None)
    if group is not None:
        from . import ulysses as U

        P = torch.distributed.get_world_size(group)
        if q.shape[1] % P:
            raise VB200Error(f"num_query_heads ({q.shape[1]}) must be divisible by ulysses_size ({P})")
        if P > k.shape[1]:  # KV head replication (:245-255)
            k = torch.repeat_interleave(k, P // k.shape[1], dim=1)
            v = torch.repeat_interleave(v, P // v.shape[1], dim=1)
        q, k, v = U.gather_seq_scatter_heads_qkv(q.contiguous(), k.contiguous(), v.contiguous(), seq_dim=0, head_dim=1, group=group)
    if max_len <= 0:
        max_len = int((cu[1:] - cu[:-1]).max())
    o = flash_attn_varlen(q, k, v, cu, max_len, scaling, bool(is_causal))
    if group is not None:
        o = U.gather_heads_scatter_seq(o, head_dim=1, seq_dim=0, group=group)
    return o.unsqueeze(0), None
