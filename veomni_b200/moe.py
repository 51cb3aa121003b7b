"""Fused MoE expert computation on the sm_100a kernels (routing, permutation, tcgen05 GroupGEMM).

Mirrors the reference surface:

* ``group_gemm_same_nk`` / ``group_gemm_same_mn``  (veomni/ops/kernels/moe/_kernels/kernel/group_gemm.py:157-234, 357-397)
* ``expert_histogram`` + scatter index, ``moe_scatter`` / ``moe_gather``  (_kernels/kernel/moe.py)
* ``fused_moe_forward(num_experts, routing_weights, selected_experts, hidden_states, fc1_1_weight,
  fc1_2_weight, fc2_weight, fc1_1_2_weight=None)`` — the raw callable VeOmni stores in
  ``veomni.ops.kernels.moe._fused_moe_forward`` (ops/kernels/moe/__init__.py:30-59) with the op order
  of ``MergedFc1TritonFusedMoeExpertFunction`` (ops/kernels/moe/group_gemm.py:269-444): the routing
  weight multiplies the fc1 activation *before* fc2 in the non-EP path.
"""

from __future__ import annotations

import torch

from . import _lib, prof
from . import functional as F
from ._lib import VB200Error, check, stream_ptr

BF = torch.bfloat16


def _req(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != BF:
        raise VB200Error(f"{name}: expected a CUDA bfloat16 tensor (no CPU fallback)")
    return t.contiguous()


# ---------------------------------------------------------------------------------------------------
# routing
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def moe_route(expert_index: torch.Tensor, num_experts: int):
    """Returns (splits int32 [E], cumsum int32 [E] inclusive, scatter_index int32 like expert_index)."""
    if not expert_index.is_cuda or expert_index.dtype not in (torch.int64, torch.int32):
        raise VB200Error("moe_route: expert_index must be a CUDA int32/int64 tensor")
    idx = expert_index.contiguous()
    n = idx.numel()
    dev = idx.device
    splits = torch.empty(num_experts, dtype=torch.int32, device=dev)
    cumsum = torch.empty(num_experts, dtype=torch.int32, device=dev)
    sidx = torch.empty(idx.shape, dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws = torch.empty(max(1, lib.vb200_moe_route_workspace(n, num_experts)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.vb200_moe_route(idx.data_ptr(), 1 if idx.dtype == torch.int64 else 0, n, num_experts, splits.data_ptr(),
                                  cumsum.data_ptr(), sidx.data_ptr(), ws.data_ptr(), stream_ptr()), "vb200_moe_route")
    return splits, cumsum, sidx


def _scatter_raw(x, sidx, w=None):
    T, K = sidx.shape
    H = x.shape[-1]
    out = torch.empty(T * K, H, dtype=x.dtype, device=x.device)
    w_out = torch.empty(T * K, 1, dtype=w.dtype, device=x.device) if w is not None else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        check(lib.vb200_moe_scatter(x.data_ptr(), sidx.data_ptr(), out.data_ptr(), w.data_ptr() if w is not None else None,
                                    w_out.data_ptr() if w is not None else None, T, K, H, stream_ptr()), "vb200_moe_scatter")
    return out, w_out


def _gather_raw(x, sidx, w=None):
    T, K = sidx.shape
    H = x.shape[-1]
    out = torch.empty(T, H, dtype=x.dtype, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        check(lib.vb200_moe_gather(x.data_ptr(), sidx.data_ptr(), w.data_ptr() if w is not None else None, out.data_ptr(),
                                   T, K, H, stream_ptr()), "vb200_moe_gather")
    return out


class _MoeScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sidx):
        ctx.save_for_backward(sidx)
        return _scatter_raw(_req(x, "moe_scatter"), sidx)[0]

    @staticmethod
    def backward(ctx, g):
        (sidx,) = ctx.saved_tensors
        return _gather_raw(g.contiguous(), sidx), None


class _MoeGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sidx):
        ctx.save_for_backward(sidx)
        return _gather_raw(_req(x, "moe_gather"), sidx)

    @staticmethod
    def backward(ctx, g):
        (sidx,) = ctx.saved_tensors
        return _scatter_raw(g.contiguous(), sidx)[0], None


def moe_scatter(x: torch.Tensor, scatter_index: torch.Tensor) -> torch.Tensor:
    """``out[scatter_index[t,k]] = x[t]`` (reference: _kernels/kernel/moe.py:253-333)."""
    return _MoeScatter.apply(x, scatter_index)


def moe_gather(x: torch.Tensor, scatter_index: torch.Tensor) -> torch.Tensor:
    """``out[t] = sum_k x[scatter_index[t,k]]`` in fp32 (reference: _kernels/kernel/moe.py:87-159)."""
    return _MoeGather.apply(x, scatter_index)


# ---------------------------------------------------------------------------------------------------
# GroupGEMM
# ---------------------------------------------------------------------------------------------------
def _cumsum32(c: torch.Tensor) -> torch.Tensor:
    return c if c.dtype == torch.int32 else c.to(torch.int32)


# 0: library default (tokens on the MMA N side for the ragged-M modes); 1: classic (tokens on M); 2: swapped. Tests compare them.
GG_VARIANT = 0


def _gg(mode: int, a, b, c, cumsum, G, rows, m, n, k):
    lib = _lib.load()
    mode = mode | (GG_VARIANT << 8)
    flops = 2.0 * rows * (m * n if (mode & 0xff) == 2 else n * k)
    with torch.cuda.device(a.device), prof.span("group_gemm", flops):
        check(lib.vb200_group_gemm(mode, a.data_ptr(), b.data_ptr(), c.data_ptr(), cumsum.data_ptr(), G, rows, m, n, k,
                                   stream_ptr()), "vb200_group_gemm")
    return c


def group_gemm_same_nk(a: torch.Tensor, b: torch.Tensor, cumsum_M: torch.Tensor, max_M: int | None = None,
                       transpose_a: bool = False, transpose_b: bool = False) -> torch.Tensor:
    """Ragged-M GroupGEMM (no autograd).  Same arguments as the reference's ``group_gemm_same_nk``."""
    if transpose_a:
        raise VB200Error("group_gemm_same_nk: transpose_a is not supported (neither does the reference test it)")
    a, b = _req(a, "a"), _req(b, "b")
    G = b.shape[0]
    if transpose_b:
        N, K = b.shape[1], b.shape[2]
    else:
        K, N = b.shape[1], b.shape[2]
    if a.shape[1] != K or cumsum_M.numel() != G:
        raise VB200Error("group_gemm_same_nk: shape mismatch")
    c = torch.empty(a.shape[0], N, dtype=a.dtype, device=a.device)
    if a.shape[0] == 0:
        return c
    return _gg(0 if transpose_b else 1, a, b, c, _cumsum32(cumsum_M), G, a.shape[0], 0, N, K)


def group_gemm_same_mn(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, cumsum_K: torch.Tensor, max_K: int | None = None,
                       transpose_a: bool = True, transpose_b: bool = False) -> torch.Tensor:
    """Ragged-K GroupGEMM ``c[g] = a[rows g]^T @ b[rows g]`` written into ``c`` [G, M, N]."""
    if not transpose_a or transpose_b:
        raise VB200Error("group_gemm_same_mn: only transpose_a=True, transpose_b=False (the wgrad form) is supported")
    a, b = _req(a, "a"), _req(b, "b")
    G, M, N = c.shape
    if a.shape[1] != M or b.shape[1] != N or a.shape[0] != b.shape[0] or not c.is_contiguous():
        raise VB200Error("group_gemm_same_mn: shape mismatch")
    if a.shape[0] == 0:
        return c.zero_()
    return _gg(2, a, b, c, _cumsum32(cumsum_K), G, a.shape[0], M, N, 0)


# ---------------------------------------------------------------------------------------------------
# fused MoE (non-EP), merged fc1
# ---------------------------------------------------------------------------------------------------
class _FusedMoeMerged(torch.autograd.Function):
    """MergedFc1TritonFusedMoeExpertFunction (ops/kernels/moe/group_gemm.py:269-444) on sm_100a kernels."""

    @staticmethod
    def forward(ctx, num_experts, gate_weights, expert_index, hidden_states, fc1_1_2_weight, fc2_weight):
        hs = _req(hidden_states.reshape(-1, hidden_states.shape[-1]), "hidden_states")
        gw = _req(gate_weights, "routing_weights")
        w1, w2 = _req(fc1_1_2_weight, "fc1_1_2_weight"), _req(fc2_weight, "fc2_weight")
        inter = w1.shape[1] // 2
        splits, cumsum, sidx = moe_route(expert_index, num_experts)
        scatter_output, scattered_gate_weight = _scatter_raw(hs, sidx, gw)
        fc1_output = group_gemm_same_nk(scatter_output, w1, cumsum, transpose_b=True)
        fc1_1_output, fc1_2_output = fc1_output[:, :inter], fc1_output[:, inter:]
        with torch.no_grad():
            fc1_activation = F.silu_mul(fc1_1_output, fc1_2_output)
        fc1_weighted_output = fc1_activation * scattered_gate_weight
        fc2_output = group_gemm_same_nk(fc1_weighted_output, w2, cumsum, transpose_b=True)
        output = _gather_raw(fc2_output, sidx).reshape(hidden_states.shape)
        ctx.save_for_backward(gw, w1, w2, sidx, scatter_output, cumsum, fc1_output, fc1_activation, scattered_gate_weight,
                              fc1_weighted_output)
        ctx.hs_shape = hidden_states.shape
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (gw, w1, w2, sidx, scatter_output, cumsum, fc1_output, fc1_activation, scattered_gate_weight,
         fc1_weighted_output) = ctx.saved_tensors
        inter = w1.shape[1] // 2
        G = w1.shape[0]
        go = grad_output.reshape(-1, grad_output.shape[-1]).contiguous()
        grad_fc2_output, _ = _scatter_raw(go, sidx)
        grad_fc1_weighted = group_gemm_same_nk(grad_fc2_output, w2, cumsum, transpose_b=False)
        grad_w2 = None
        if ctx.needs_input_grad[5]:
            grad_w2 = torch.empty_like(w2)
            group_gemm_same_mn(grad_fc2_output, fc1_weighted_output, grad_w2, cumsum)
        grad_fc1_activation = grad_fc1_weighted * scattered_gate_weight
        grad_scattered_gw = torch.sum(fc1_activation * grad_fc1_weighted, dim=-1)
        grad_gate_weight = grad_scattered_gw[sidx.flatten().long()].reshape(gw.shape)
        # silu backward straight into the two halves of the merged [T*K, 2I] gradient
        grad_fc1_output = torch.empty_like(fc1_output)
        rows = fc1_output.shape[0]
        lib = _lib.load()
        gfa = grad_fc1_activation.contiguous()
        with torch.cuda.device(gfa.device):
            check(lib.vb200_swiglu_bwd(gfa.data_ptr(), fc1_output.data_ptr(), fc1_output.data_ptr() + inter * 2,
                                       grad_fc1_output.data_ptr(), grad_fc1_output.data_ptr() + inter * 2, rows, inter,
                                       2 * inter, inter, 2 * inter, stream_ptr()), "vb200_swiglu_bwd")
        grad_scatter_output = group_gemm_same_nk(grad_fc1_output, w1, cumsum, transpose_b=False)
        grad_w1 = None
        if ctx.needs_input_grad[4]:
            grad_w1 = torch.empty_like(w1)
            group_gemm_same_mn(grad_fc1_output, scatter_output, grad_w1, cumsum)
        grad_hidden = _gather_raw(grad_scatter_output, sidx).reshape(ctx.hs_shape)
        return None, grad_gate_weight, None, grad_hidden, grad_w1, grad_w2


def fused_moe_forward(num_experts: int, routing_weights: torch.Tensor, selected_experts: torch.Tensor,
                      hidden_states: torch.Tensor, fc1_1_weight: torch.Tensor | None, fc1_2_weight: torch.Tensor | None,
                      fc2_weight: torch.Tensor, fc1_1_2_weight: torch.Tensor | None = None) -> torch.Tensor:
    """Same contract as the reference's raw fused-MoE callable (ops/kernels/moe/group_gemm.py:447-549)."""
    if fc1_1_2_weight is None:
        if fc1_1_weight is None or fc1_2_weight is None:
            raise ValueError("Split fc1 mode requires both fc1_1_weight and fc1_2_weight.")
        fc1_1_2_weight = torch.cat([fc1_1_weight, fc1_2_weight], dim=1)
    elif fc1_1_weight is not None or fc1_2_weight is not None:
        raise ValueError("Provide either split fc1 weights or merged fc1_1_2_weight, not both.")
    ep = _ep_state()
    if ep is not None:
        from . import ep as _ep

        return _ep.ep_fused_moe_forward(ep, num_experts, routing_weights, selected_experts, hidden_states,
                                        fc1_1_2_weight, fc2_weight)
    return _FusedMoeMerged.apply(num_experts, routing_weights, selected_experts, hidden_states, fc1_1_2_weight, fc2_weight)


_EP = None


def set_ep_group(state) -> None:
    """Install (or clear with None) the expert-parallel context created by ``veomni_b200.ep.EPContext``."""
    global _EP
    _EP = state


def _ep_state():
    return _EP


def moe_experts_forward(module, hidden_states: torch.Tensor, top_k_index: torch.Tensor, top_k_weights: torch.Tensor):
    """Drop-in for OpSlot("moe_experts", "standard"): ``f(self, hidden_states[T,H], top_k_index[T,K], top_k_weights[T,K])``
    with merged ``gate_up_proj [E,2I,H]`` / ``down_proj [E,H,I]`` (ops/kernels/moe/__init__.py:133-143)."""
    return fused_moe_forward(module.num_experts, top_k_weights.to(hidden_states.dtype), top_k_index, hidden_states,
                             None, None, module.down_proj, fc1_1_2_weight=module.gate_up_proj)
