"""Async Ulysses: q/k/v projections interleaved with their sequence-parallel all-to-all, and the reverse for o_proj.

Mirror of veomni/distributed/sequence_parallel/async_ulysses.py (``AsyncUlyssesQKVProjection`` :48-333,
``AsyncUlyssesOutputProjection`` :336-466, functional faces ``async_ulysses_qkv_projection`` /
``async_ulysses_output_projection`` :469-503) — the attention front/back end Qwen3-VL and Wan call
(models/transformers/qwen3_vl/generated/patched_modeling_qwen3_vl_gpu.py:175-232). Same argument names, same results as
the synchronous path (projection -> ``gather_seq_scatter_heads`` -> q/k norm), which is the parity target the reference
itself uses (tests/parallel/ulysses/test_async_ulysses.py:112-115).

What differs is the machinery:
* the exchange is the NVLink pull kernel (``ulysses.all_to_all_many``), launched on a dedicated communication stream
  right after the GEMM that produced its input, so q's exchange runs under the k projection, k's under the v projection
  (forward), and the gradient exchanges under the weight / input gradient GEMMs (backward) — the reference gets the same
  overlap from ``dist.all_to_all_single(async_op=True)``;
* q/k normalisation after the gather is this package's RMSNorm kernel (the reference needs apex's
  ``fused_layer_norm_cuda``, which is not in the image) — ``norm_type`` "rmsnorm" or None; "layernorm" raises.
"""

from __future__ import annotations

from typing import Any

import torch
import torch.nn.functional as Fnn

from . import _lib
from ._lib import VB200Error, check, stream_ptr
from .ulysses import all_to_all_many

_comm_streams: dict[int, torch.cuda.Stream] = {}


def _comm_stream(dev: torch.device) -> torch.cuda.Stream:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _comm_streams:
        _comm_streams[idx] = torch.cuda.Stream(device=dev)
    return _comm_streams[idx]


class _Pending:
    """An exchange in flight on the communication stream; ``wait()`` makes the current stream depend on it."""

    def __init__(self, out: torch.Tensor, ev: torch.cuda.Event):
        self.out, self.ev = out, ev

    def wait(self) -> torch.Tensor:
        torch.cuda.current_stream().wait_event(self.ev)
        self.out.record_stream(torch.cuda.current_stream())
        return self.out


def _a2a_async(x: torch.Tensor, scatter_dim: int, gather_dim: int, group) -> _Pending:
    cur = torch.cuda.current_stream()
    side = _comm_stream(x.device)
    ready = torch.cuda.Event()
    ready.record(cur)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        x.record_stream(side)
        out = all_to_all_many([x.contiguous()], scatter_dim, gather_dim, group)[0]
        done = torch.cuda.Event()
        done.record(side)
    return _Pending(out, done)


def _pad_seq(x: torch.Tensor, dim: int, world: int) -> torch.Tensor:
    n = x.size(dim)
    if n % world == 0:
        return x
    pad = list(x.shape)
    pad[dim] = world - n % world
    return torch.cat([x, x.new_zeros(pad)], dim=dim)


def _unpad(x: torch.Tensor, dim: int, unpadded: int) -> torch.Tensor:
    return x.narrow(dim, 0, unpadded) if unpadded and x.size(dim) != unpadded else x


def _rms_fwd(x: torch.Tensor, w: torch.Tensor, eps: float):
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    y = torch.empty_like(x2)
    rstd = torch.empty(x2.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().vb200_rmsnorm_fwd(x2.data_ptr(), w.data_ptr(), y.data_ptr(), rstd.data_ptr(), x2.shape[0], x2.shape[1],
                                            float(eps), stream_ptr()), "vb200_rmsnorm_fwd")
    return y.view(x.shape), rstd


def _rms_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor):
    lib = _lib.load()
    x2, dy2 = x.reshape(-1, x.shape[-1]).contiguous(), dy.reshape(-1, dy.shape[-1]).contiguous()
    dx = torch.empty_like(x2)
    parts = torch.empty(max(1, lib.vb200_rmsnorm_bwd_partials(x2.shape[0], x2.shape[1])), x2.shape[1], dtype=torch.float32, device=x.device)
    dw = torch.empty(x2.shape[1], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.vb200_rmsnorm_bwd(dy2.data_ptr(), x2.data_ptr(), w.data_ptr(), rstd.data_ptr(), dx.data_ptr(), parts.data_ptr(),
                                    dw.data_ptr(), x2.shape[0], x2.shape[1], stream_ptr()), "vb200_rmsnorm_bwd")
    return dx.view(x.shape), dw.to(w.dtype)


class AsyncUlyssesQKVProjection(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, hidden_states, seq_dimension, head_dimension, q_weight, q_bias, k_weight, k_bias, v_weight, v_bias,
                norm_type, norm_q_weight, norm_q_bias, norm_k_weight, norm_k_bias, normalized_shape, eps, unpadded_dim_size,
                head_dim, group):
        if norm_type not in (None, "rmsnorm"):
            raise NotImplementedError(f"{norm_type} is not supported in async-ulysses now!")
        world = torch.distributed.get_world_size(group)
        nq, nkv = q_weight.shape[0] // head_dim, k_weight.shape[0] // head_dim
        B = hidden_states.shape[0]
        if nq % world:
            raise VB200Error(f"num_query_heads ({nq}) must be divisible by ulysses_size ({world})")
        rep = 1
        if world > nkv:
            if world % nkv:
                raise VB200Error(f"ulysses_size ({world}) must be divisible by num_key_value_heads ({nkv})")
            rep = world // nkv
        q = Fnn.linear(hidden_states, q_weight, q_bias).view(B, -1, nq, head_dim)
        q_p = _a2a_async(q, head_dimension, seq_dimension, group)          # runs under the k projection
        k = Fnn.linear(hidden_states, k_weight, k_bias).view(B, -1, nkv, head_dim)
        if rep > 1:
            k = torch.repeat_interleave(k, rep, dim=2)
        k_p = _a2a_async(k, head_dimension, seq_dimension, group)          # runs under the v projection
        v = Fnn.linear(hidden_states, v_weight, v_bias).view(B, -1, nkv, head_dim)
        if rep > 1:
            v = torch.repeat_interleave(v, rep, dim=2)
        v_p = _a2a_async(v, head_dimension, seq_dimension, group)
        q = _unpad(q_p.wait(), seq_dimension, unpadded_dim_size).contiguous()
        k = _unpad(k_p.wait(), seq_dimension, unpadded_dim_size).contiguous()
        rq = rk = None
        oq, ok = q, k
        if norm_type == "rmsnorm":
            oq, rq = _rms_fwd(q, norm_q_weight.contiguous(), eps)         # under v's exchange
            ok, rk = _rms_fwd(k, norm_k_weight.contiguous(), eps)
        v = _unpad(v_p.wait(), seq_dimension, unpadded_dim_size)
        ctx.group, ctx.seq_dim, ctx.head_dim_ix, ctx.norm_type, ctx.rep, ctx.nkv, ctx.world = (
            group, seq_dimension, head_dimension, norm_type, rep, nkv, world)
        ctx.save_for_backward(hidden_states, q_weight, q_bias, k_weight, k_bias, v_weight, v_bias, q, norm_q_weight, rq, k,
                              norm_k_weight, rk)
        return oq, ok, v

    @staticmethod
    def backward(ctx: Any, gq, gk, gv):
        (hs, qw, qb, kw, kb, vw, vb, q, nqw, rq, k, nkw, rk) = ctx.saved_tensors
        group, sd, hd, rep, nkv, world = ctx.group, ctx.seq_dim, ctx.head_dim_ix, ctx.rep, ctx.nkv, ctx.world
        B = hs.shape[0]
        hs2 = hs.reshape(-1, hs.shape[-1])

        def back(g):  # gradient of (exchange -> unpad): pad the sequence back, reverse exchange
            return _a2a_async(_pad_seq(g.contiguous(), sd, world), sd, hd, group)

        def fold(g):  # gradient of repeat_interleave over the kv heads
            return g.reshape(g.shape[0], g.shape[1], nkv, rep, g.shape[-1]).sum(dim=3) if rep > 1 else g

        gv_p = back(gv)                                                     # runs under the norm backward
        g_nq = g_nk = None
        if ctx.norm_type == "rmsnorm":
            gq, g_nq = _rms_bwd(gq, q, nqw.contiguous(), rq)
            gk, g_nk = _rms_bwd(gk, k, nkw.contiguous(), rk)
        gvl = fold(gv_p.wait())
        gk_p = back(gk)                                                     # runs under the v-projection gradients
        gv2 = gvl.reshape(-1, vw.shape[0])
        g_in = gv2 @ vw
        g_vw = gv2.t() @ hs2
        g_vb = gv2.sum(0) if vb is not None else None
        gkl = fold(gk_p.wait())
        gq_p = back(gq)                                                     # runs under the k-projection gradients
        gk2 = gkl.reshape(-1, kw.shape[0])
        g_in = g_in + gk2 @ kw
        g_kw = gk2.t() @ hs2
        g_kb = gk2.sum(0) if kb is not None else None
        gq2 = gq_p.wait().reshape(-1, qw.shape[0])
        g_in = g_in + gq2 @ qw
        g_qw = gq2.t() @ hs2
        g_qb = gq2.sum(0) if qb is not None else None
        return (g_in.view(hs.shape), None, None, g_qw, g_qb, g_kw, g_kb, g_vw, g_vb, None, g_nq, None, g_nk, None, None, None,
                None, None, None)


class AsyncUlyssesOutputProjection(torch.autograd.Function):
    @staticmethod
    def forward(ctx: Any, hidden_states, seq_dimension, head_dimension, proj_weight, proj_bias, unpadded_dim_size, group):
        world = torch.distributed.get_world_size(group)
        x = _pad_seq(hidden_states, seq_dimension, world)
        x = all_to_all_many([x.contiguous()], seq_dimension, head_dimension, group)[0]  # nothing to overlap with: o_proj needs it
        ctx.heads, ctx.hdim = x.shape[head_dimension], x.shape[-1]
        x = x.reshape(x.shape[0], x.shape[1], -1)
        ctx.group, ctx.seq_dim, ctx.head_dim_ix, ctx.unpadded = group, seq_dimension, head_dimension, unpadded_dim_size
        ctx.save_for_backward(x, proj_weight, proj_bias)
        return Fnn.linear(x, proj_weight, proj_bias)

    @staticmethod
    def backward(ctx: Any, g):
        x, w, b = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        go = (g2 @ w).view(g.shape[0], -1, ctx.heads, ctx.hdim)
        go_p = _a2a_async(go, ctx.head_dim_ix, ctx.seq_dim, ctx.group)      # runs under the weight gradient
        g_w = g2.t() @ x.reshape(-1, x.shape[-1])
        g_b = g2.sum(0) if b is not None else None
        go = _unpad(go_p.wait(), ctx.seq_dim, ctx.unpadded)
        return go, None, None, g_w, g_b, None, None


def async_ulysses_qkv_projection(hidden_states, seq_dimension, head_dimension, q_weight, q_bias, k_weight, k_bias, v_weight,
                                 v_bias, norm_type, norm_q_weight, norm_q_bias, norm_k_weight, norm_k_bias, normalized_shape, eps,
                                 unpadded_dim_size, head_dim, group=None):
    """Reference: async_ulysses.py:469-492 (same keyword arguments)."""
    return AsyncUlyssesQKVProjection.apply(hidden_states, seq_dimension, head_dimension, q_weight, q_bias, k_weight, k_bias,
                                           v_weight, v_bias, norm_type, norm_q_weight, norm_q_bias, norm_k_weight, norm_k_bias,
                                           normalized_shape, eps, unpadded_dim_size, head_dim, group)


def async_ulysses_output_projection(hidden_states, seq_dimension, head_dimension, proj_weight, proj_bias, unpadded_dim_size,
                                    group=None):
    """Reference: async_ulysses.py:495-503."""
    return AsyncUlyssesOutputProjection.apply(hidden_states, seq_dimension, head_dimension, proj_weight, proj_bias,
                                              unpadded_dim_size, group)


def install() -> None:
    """Route VeOmni's async-Ulysses entry points through this module (no-op if VeOmni is not importable)."""
    try:
        from veomni.distributed.sequence_parallel import async_ulysses as ref
    except ImportError:
        return
    ref.async_ulysses_qkv_projection = async_ulysses_qkv_projection
    ref.async_ulysses_output_projection = async_ulysses_output_projection
