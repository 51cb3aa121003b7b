"""Autograd-aware Python faces of the memory-bound sm_100a kernels.

Signatures mirror the callables VeOmni binds into its ``OpSlot``s:

* ``rms_norm(hidden_states, weight, eps)``           — OpSlot("rms_norm", "standard")
* ``apply_rotary_pos_emb(q, k, cos, sin, ...)``      — OpSlot("rotary_pos_emb", "full")
* ``swiglu_mlp(module, x)``                          — OpSlot("swiglu_mlp", "standard")

plus ``qknorm_rope`` (q/k head RMSNorm + RoPE in one pass), a fusion the reference cannot express
through its slots but that sits on the same call sites
(veomni/models/transformers/qwen3/generated/patched_modeling_qwen3_gpu.py:305-310).

All ops require CUDA bf16 tensors and raise :class:`veomni_b200._lib.VB200Error` otherwise —
there is no eager fallback.
"""

from __future__ import annotations

import torch

from . import _lib
from ._lib import VB200Error, check, stream_ptr


def _need_cuda_bf16(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if not t.is_cuda:
            raise VB200Error("veomni_b200 ops run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.bfloat16:
            raise VB200Error(f"veomni_b200 ops expect bfloat16 tensors, got {t.dtype}")


def _aligned(t: torch.Tensor) -> torch.Tensor:
    if t.data_ptr() % 16:
        return t.clone(memory_format=torch.contiguous_format)
    return t


# ----------------------------------------------------------------------------------------------
# RMSNorm
# ----------------------------------------------------------------------------------------------
class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, eps: float):
        _need_cuda_bf16(x)
        cols = x.shape[-1]
        x2 = _aligned(x.reshape(-1, cols).contiguous())
        w = _aligned(weight.detach().to(torch.bfloat16).contiguous())
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            check(
                lib.vb200_rmsnorm_fwd(x2.data_ptr(), w.data_ptr(), y.data_ptr(), rstd.data_ptr(), rows, cols,
                                      float(eps), stream_ptr()),
                "vb200_rmsnorm_fwd",
            )
        ctx.save_for_backward(x2, w, rstd)
        ctx.x_shape = x.shape
        ctx.w_dtype = weight.dtype
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        x2, w, rstd = ctx.saved_tensors
        rows, cols = x2.shape
        dy2 = _aligned(dy.reshape(rows, cols).contiguous())
        lib = _lib.load()
        dx = torch.empty_like(x2)
        nparts = lib.vb200_rmsnorm_bwd_partials(rows, cols)
        partial = torch.empty(max(nparts, 1), cols, dtype=torch.float32, device=x2.device)
        dw = torch.empty(cols, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            check(
                lib.vb200_rmsnorm_bwd(dy2.data_ptr(), x2.data_ptr(), w.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                      partial.data_ptr(), dw.data_ptr(), rows, cols, stream_ptr()),
                "vb200_rmsnorm_bwd",
            )
        return dx.view(ctx.x_shape), dw.to(ctx.w_dtype), None


def rms_norm(hidden_states: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """``weight * (x * rsqrt(mean(x^2) + eps)).to(bf16)`` — drop-in for the ``rms_norm`` OpSlot."""
    return _RMSNorm.apply(hidden_states, weight, eps)


class _AddRMSNorm(torch.autograd.Function):
    """(y, h) = (RMSNorm(x + residual) * w, x + residual) in one pass (rmsnorm.cu add_rmsnorm_fwd_kernel / rmsnorm_bwd_ring_kernel<.., ADD>); on by default in the decoder layer."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float):
        _need_cuda_bf16(x)
        _need_cuda_bf16(residual)
        if x.shape != residual.shape:
            raise VB200Error("fused_add_rms_norm: x and residual must have the same shape")
        cols = x.shape[-1]
        x2 = _aligned(x.reshape(-1, cols).contiguous())
        r2 = _aligned(residual.reshape(-1, cols).contiguous())
        w = _aligned(weight.detach().to(torch.bfloat16).contiguous())
        rows = x2.shape[0]
        h, y = torch.empty_like(x2), torch.empty_like(x2)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(_lib.load().vb200_add_rmsnorm_fwd(x2.data_ptr(), r2.data_ptr(), w.data_ptr(), h.data_ptr(), y.data_ptr(),
                                                    rstd.data_ptr(), rows, cols, float(eps), stream_ptr()), "vb200_add_rmsnorm_fwd")
        ctx.save_for_backward(h, w, rstd)
        ctx.x_shape, ctx.w_dtype = x.shape, weight.dtype
        return y.view(x.shape), h.view(x.shape)

    @staticmethod
    def backward(ctx, dy: torch.Tensor, dh: torch.Tensor | None):
        h, w, rstd = ctx.saved_tensors
        rows, cols = h.shape
        dy2 = _aligned(dy.reshape(rows, cols).contiguous())
        dh2 = _aligned(dh.reshape(rows, cols).contiguous()) if dh is not None else None
        lib = _lib.load()
        dx = torch.empty_like(h)
        partial = torch.empty(max(lib.vb200_rmsnorm_bwd_partials(rows, cols), 1), cols, dtype=torch.float32, device=h.device)
        dw = torch.empty(cols, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            check(lib.vb200_rmsnorm_bwd_add(dy2.data_ptr(), h.data_ptr(), w.data_ptr(), rstd.data_ptr(),
                                            dh2.data_ptr() if dh2 is not None else None, dx.data_ptr(), partial.data_ptr(),
                                            dw.data_ptr(), rows, cols, stream_ptr()), "vb200_rmsnorm_bwd_add")
        dx = dx.view(ctx.x_shape)
        return dx, dx, dw.to(ctx.w_dtype), None


FUSED_ADD_NORM_COLS = (1024, 2048, 4096, 5120, 8192)  # hidden sizes vb200_add_rmsnorm_fwd is instantiated for


def fused_add_rms_norm(x: torch.Tensor, residual: torch.Tensor, weight: torch.Tensor, eps: float):
    """``h = residual + x`` (bf16) and ``rms_norm(h, weight, eps)`` in one kernel; returns ``(normed, h)``.
    Replaces the `hidden_states = residual + hidden_states` + RMSNorm pair of the decoder layer
    (patched_modeling_qwen3_gpu.py:369-375). Hidden sizes: ``FUSED_ADD_NORM_COLS``."""
    return _AddRMSNorm.apply(x, residual, weight, eps)


# ----------------------------------------------------------------------------------------------
# RoPE
# ----------------------------------------------------------------------------------------------
def _rope_launch(q_in, q_out, k_in, k_out, cos, sin, inverse: bool) -> None:
    """q/k: [S, H, D]-addressable views (stride(-1) == 1); cos/sin: [S, D] contiguous."""
    S, Hq, D = q_in.shape
    Hk = k_in.shape[1]
    lib = _lib.load()
    with torch.cuda.device(q_in.device):
        check(
            lib.vb200_rope(
                q_in.data_ptr(), q_out.data_ptr(), k_in.data_ptr(), k_out.data_ptr(), cos.data_ptr(),
                sin.data_ptr(), S, Hq, Hk, D,
                q_in.stride(0), q_in.stride(1), k_in.stride(0), k_in.stride(1),
                q_out.stride(0), q_out.stride(1), k_out.stride(0), k_out.stride(1),
                1 if inverse else 0, stream_ptr(),
            ),
            "vb200_rope",
        )


def _shd_view(t: torch.Tensor, b: int) -> torch.Tensor:
    """[B, H, S, D] tensor -> [S, H, D] view of batch element b (no copy when stride(-1) == 1)."""
    v = t[b].transpose(0, 1)
    if v.stride(-1) != 1 or v.data_ptr() % 16 or v.stride(0) % 8 or v.stride(1) % 8:
        v = v.contiguous()
    return v


class _RoPE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, cos, sin):
        # q: [B, Hq, S, D], k: [B, Hk, S, D] (any strides), cos/sin: [B|1, S, D]
        _need_cuda_bf16(q, k)
        B, Hq, S, D = q.shape
        Hk = k.shape[1]
        cos = cos.to(torch.bfloat16).contiguous()
        sin = sin.to(torch.bfloat16).contiguous()
        q_out = torch.empty(B, S, Hq, D, dtype=q.dtype, device=q.device)
        k_out = torch.empty(B, S, Hk, D, dtype=k.dtype, device=k.device)
        for b in range(B):
            cb = b if cos.shape[0] > 1 else 0
            _rope_launch(_shd_view(q, b), q_out[b], _shd_view(k, b), k_out[b], cos[cb], sin[cb], False)
        ctx.save_for_backward(cos, sin)
        return q_out.transpose(1, 2), k_out.transpose(1, 2)

    @staticmethod
    def backward(ctx, dq, dk):
        cos, sin = ctx.saved_tensors
        B, Hq, S, D = dq.shape
        Hk = dk.shape[1]
        dq_in = torch.empty(B, S, Hq, D, dtype=dq.dtype, device=dq.device)
        dk_in = torch.empty(B, S, Hk, D, dtype=dk.dtype, device=dk.device)
        for b in range(B):
            cb = b if cos.shape[0] > 1 else 0
            _rope_launch(_shd_view(dq, b), dq_in[b], _shd_view(dk, b), dk_in[b], cos[cb], sin[cb], True)
        return dq_in.transpose(1, 2), dk_in.transpose(1, 2), None, None


def apply_rotary_pos_emb(q, k, cos, sin, position_ids=None, unsqueeze_dim: int = 1):
    """Drop-in for the ``rotary_pos_emb`` OpSlot: q ``[B,Hq,S,D]``, k ``[B,Hkv,S,D]``,
    cos/sin ``[B|1,S,D]``; returns rotated (q, k) with the same logical shapes."""
    if unsqueeze_dim != 1:
        raise VB200Error("veomni_b200 RoPE supports the [B, H, S, D] layout (unsqueeze_dim=1) only")
    return _RoPE.apply(q, k, cos, sin)


# ----------------------------------------------------------------------------------------------
# fused q/k RMSNorm + RoPE
# ----------------------------------------------------------------------------------------------
class _QKNormRope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, wq, wk, cos, sin, eps, out=None):
        # q: [T, Hq, D], k: [T, Hk, D] contiguous; cos/sin: [T, D]
        _need_cuda_bf16(q, k)
        q = _aligned(q.contiguous())
        k = _aligned(k.contiguous())
        T, Hq, D = q.shape
        Hk = k.shape[1]
        wq_b = wq.detach().to(torch.bfloat16).contiguous()
        wk_b = wk.detach().to(torch.bfloat16).contiguous()
        cos = cos.to(torch.bfloat16).contiguous()
        sin = sin.to(torch.bfloat16).contiguous()
        if out is not None:  # e.g. views of the Ulysses staging buffer: the exchange then needs no copy-in
            q_out, k_out = out
            if q_out.shape != q.shape or k_out.shape != k.shape or not (q_out.is_contiguous() and k_out.is_contiguous()):
                raise VB200Error("qknorm_rope: `out` tensors must be contiguous and shaped like q / k")
        else:
            q_out, k_out = torch.empty_like(q), torch.empty_like(k)
        rq = torch.empty(T, Hq, dtype=torch.float32, device=q.device)
        rk = torch.empty(T, Hk, dtype=torch.float32, device=q.device)
        lib = _lib.load()
        with torch.cuda.device(q.device):
            check(
                lib.vb200_qknorm_rope_fwd(q.data_ptr(), k.data_ptr(), wq_b.data_ptr(), wk_b.data_ptr(),
                                          cos.data_ptr(), sin.data_ptr(), q_out.data_ptr(), k_out.data_ptr(),
                                          rq.data_ptr(), rk.data_ptr(), T, Hq, Hk, D, float(eps), stream_ptr()),
                "vb200_qknorm_rope_fwd",
            )
        ctx.save_for_backward(q, k, wq_b, wk_b, cos, sin, rq, rk)
        ctx.w_dtypes = (wq.dtype, wk.dtype)
        return q_out, k_out

    @staticmethod
    def backward(ctx, dq, dk):
        q, k, wq_b, wk_b, cos, sin, rq, rk = ctx.saved_tensors
        T, Hq, D = q.shape
        Hk = k.shape[1]
        dq = _aligned(dq.contiguous())
        dk = _aligned(dk.contiguous())
        dq_in, dk_in = torch.empty_like(q), torch.empty_like(k)
        lib = _lib.load()
        nparts = lib.vb200_qknorm_rope_bwd_partials(T)
        partial = torch.empty(nparts, 2 * D, dtype=torch.float32, device=q.device)
        dwq = torch.empty(D, dtype=torch.float32, device=q.device)
        dwk = torch.empty(D, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            check(
                lib.vb200_qknorm_rope_bwd(dq.data_ptr(), dk.data_ptr(), q.data_ptr(), k.data_ptr(), wq_b.data_ptr(),
                                          wk_b.data_ptr(), cos.data_ptr(), sin.data_ptr(), rq.data_ptr(),
                                          rk.data_ptr(), dq_in.data_ptr(), dk_in.data_ptr(), partial.data_ptr(),
                                          dwq.data_ptr(), dwk.data_ptr(), T, Hq, Hk, D, stream_ptr()),
                "vb200_qknorm_rope_bwd",
            )
        return dq_in, dk_in, dwq.to(ctx.w_dtypes[0]), dwk.to(ctx.w_dtypes[1]), None, None, None, None


def qknorm_rope(q, k, q_norm_weight, k_norm_weight, cos, sin, eps: float, out=None):
    """Per-head RMSNorm of q ``[T,Hq,D]`` / k ``[T,Hk,D]`` followed by RoPE, one HBM pass.
    ``out=(q_out, k_out)`` writes the results into caller-provided buffers (e.g. the Ulysses send staging)."""
    return _QKNormRope.apply(q, k, q_norm_weight, k_norm_weight, cos, sin, eps, out)


# ----------------------------------------------------------------------------------------------
# SwiGLU
# ----------------------------------------------------------------------------------------------
def _rows_view(t: torch.Tensor):
    """Return (tensor, rows, cols, row_stride) for a [..., cols] tensor whose rows are equally strided."""
    cols = t.shape[-1]
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0:
        return t, t.shape[0], cols, t.stride(0)
    t = _aligned(t.contiguous())
    return t, t.numel() // cols, cols, cols


class _SiLUMul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate, up):
        _need_cuda_bf16(gate, up)
        g, rows, cols, gs = _rows_view(gate)
        u, _, _, us = _rows_view(up)
        if gs != us:
            g, u = g.contiguous(), u.contiguous()
            gs = us = cols
        out = torch.empty(gate.shape, dtype=gate.dtype, device=gate.device)
        lib = _lib.load()
        with torch.cuda.device(gate.device):
            check(lib.vb200_swiglu_fwd(g.data_ptr(), u.data_ptr(), out.data_ptr(), rows, cols, gs, cols, stream_ptr()),
                  "vb200_swiglu_fwd")
        ctx.save_for_backward(g, u)
        ctx.meta = (rows, cols, gs, gate.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        g, u = ctx.saved_tensors
        rows, cols, gs, shape = ctx.meta
        d = _aligned(dout.contiguous())
        dg = torch.empty(shape, dtype=dout.dtype, device=dout.device)
        du = torch.empty(shape, dtype=dout.dtype, device=dout.device)
        lib = _lib.load()
        with torch.cuda.device(dout.device):
            check(lib.vb200_swiglu_bwd(d.data_ptr(), g.data_ptr(), u.data_ptr(), dg.data_ptr(), du.data_ptr(), rows,
                                       cols, gs, cols, cols, stream_ptr()),
                  "vb200_swiglu_bwd")
        return dg, du


def silu_mul(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """``silu(gate) * up`` (fp32 silu rounded to bf16 before the product)."""
    return _SiLUMul.apply(gate, up)


def swiglu_mlp(module, x: torch.Tensor) -> torch.Tensor:
    """Drop-in for the ``swiglu_mlp`` OpSlot (veomni/ops/liger/__init__.py:127-130)."""
    return module.down_proj(silu_mul(module.gate_proj(x), module.up_proj(x)))


def swiglu_mlp_residual(module, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """``residual + swiglu_mlp(module, x)`` with the residual add folded into the down-projection GEMM's epilogue
    (cuBLAS beta = 1) instead of a separate elementwise kernel (patched_modeling_qwen3_gpu.py:375)."""
    act = silu_mul(module.gate_proj(x), module.up_proj(x))
    if module.down_proj.bias is not None:
        return residual + module.down_proj(act)
    shape = residual.shape
    return torch.addmm(residual.reshape(-1, shape[-1]), act.reshape(-1, act.shape[-1]), module.down_proj.weight.t()).view(shape)
