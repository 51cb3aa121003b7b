"""Expert-parallel MoE dispatch / combine over NVLink peer memory.

Mirrors the EP branch of ``group_gemm_fused_moe_forward`` (veomni/ops/kernels/moe/group_gemm.py:468-524):
``preprocess`` -> ``token_pre_all2all`` -> ``EPMergedFc1GroupGemm`` -> ``tokens_post_all2all``
(veomni/distributed/moe/moe_layer.py:30-137, 307-441; moe_utils.py:19-99; comm.py:20-54), with the same
results and the same semantics (routing weight applied AFTER the return exchange, fp32 accumulation over
the top-k, expert-major / source-rank-minor order of the received tokens).

What differs is the machinery.  The reference permutes with masked_select/index_select, calls NCCL
``all_to_all_single`` with split sizes obtained through two device->host syncs, and regroups with
split+cat.  Here:
  1. the routing kernel (moe_route.cu) produces the expert-sorted order in one pass and the row scatter
     writes the permuted tokens straight into this rank's symmetric send buffer;
  2. the per-expert counts of all ranks are all-gathered through the symmetric region (512 B / rank);
     ONE device->host read of that [EP, E] matrix sizes the receive tensor (the reference needs two);
  3. one pull kernel (vb200_chunk_pull) copies, from every source rank, each (source, local expert)
     block to its final position — permute + all-to-all + sort_chunks fused, no intermediate copies;
  4. the return path pulls every (expert, source) block back from the owner's symmetric output buffer and
     the weighted top-k reduction (moe_gather with weights) replaces unpermute's scatter_add.
Backward uses the same two block lists in the opposite roles.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _lib, prof
from . import functional as F
from ._lib import VB200Error, check, stream_ptr
from .moe import _gather_raw, group_gemm_same_mn, group_gemm_same_nk, moe_route
from .symm import SymmetricMemory, get_symmetric_memory

CH_EP_COUNTS = 3
CH_EP_DISPATCH = 4
CH_EP_COMBINE = 5
BF = torch.bfloat16


class EPContext:
    """Expert-parallel group state: symmetric region + persistent staging buffers."""

    def __init__(self, group: dist.ProcessGroup | None = None, symm: SymmetricMemory | None = None, num_ctas: int = 32):
        self.group = group if group is not None else dist.group.WORLD
        self.symm = symm if symm is not None else get_symmetric_memory(self.group)
        self.ep_size = self.symm.world
        self.rank = self.symm.rank
        self.num_ctas = num_ctas
        self._bufs: dict[str, torch.Tensor] = {}
        # host-side routing statistics (rows this rank received per exchange): load imbalance is what sizes the staging
        # buffers and the GroupGEMM tail, so the bench line reports it
        self.stats = {"calls": 0, "min_recv": None, "max_recv": 0, "sum_recv": 0}

    def staging(self, tag: str, nbytes: int) -> torch.Tensor:
        """ONE persistent symmetric buffer per tag, grown geometrically: receive sizes depend on the routing and differ on
        nearly every layer and step, so a buffer per (tag, size) would leak the arena. A larger buffer replaces the old one
        (whose block returns to the arena; every user runs on the caller's stream and the pull kernels only retire after
        all peers finished reading, so stream order makes the reuse safe). Each collective publishes the offset of the
        buffer it exposes, so ranks growing at different times is fine."""
        nbytes = (max(int(nbytes), 1) + 255) // 256 * 256
        cur = self._bufs.get(tag)
        if cur is None or cur.numel() < nbytes:
            want = nbytes if cur is None else max(nbytes, cur.numel() * 3 // 2)
            want = (want + 255) // 256 * 256
            self._bufs.pop(tag, None)
            del cur
            self._bufs[tag] = self.symm.empty((want,), torch.uint8, arena="misc")
        return self._bufs[tag]


@dataclass
class EPPlan:
    """Everything derived from the routing of one MoE layer call (no autograd state)."""

    ctx: EPContext
    T: int
    K: int
    H: int
    sidx: torch.Tensor                # [T, K] int32: row of slot (t,k) in the local expert-major order
    counts: torch.Tensor              # [EP, E] int64 on host: tokens rank s sends to expert e
    input_splits: list                # [EP] rows this rank sends to each rank      (moe_layer.py:41)
    output_splits: list               # [EP] rows this rank receives from each rank (moe_layer.py:58)
    total_recv: int
    cumsum_local: torch.Tensor        # [E/EP] int32 device: inclusive prefix of rows per local expert
    fwd_chunks: torch.Tensor          # device ChunkDesc list: owner pulls (source, local expert) blocks
    n_fwd: int
    bwd_chunks: torch.Tensor          # device ChunkDesc list: source pulls (expert, me) blocks back
    n_bwd: int


def _chunk_tensor(chunks: list[tuple[int, int, int, int]], dev) -> torch.Tensor:
    """[(peer, src_off, dst_off, bytes)] -> device int64 [n, 4] laid out as the C struct ChunkDesc."""
    if not chunks:
        return torch.zeros(1, 4, dtype=torch.int64, device=dev)
    rows = [[s, d, b, p & 0xFFFFFFFF] for (p, s, d, b) in chunks]  # {src_off, dst_off, bytes, peer | pad<<32}
    return torch.tensor(rows, dtype=torch.int64).to(dev, non_blocking=True)


@torch.no_grad()
def make_plan(ctx: EPContext, selected_experts: torch.Tensor, num_experts: int, hidden: int) -> EPPlan:
    """Routing + counts exchange + block lists (reference: preprocess, moe_layer.py:30-69)."""
    ep, r = ctx.ep_size, ctx.rank
    if num_experts % ep:
        raise VB200Error("num_experts must be divisible by the EP group size")
    el = num_experts // ep
    T, K = selected_experts.shape
    dev = selected_experts.device
    splits, _cumsum, sidx = moe_route(selected_experts, num_experts)
    # all-gather the per-expert counts through the symmetric region
    cbuf = ctx.staging("counts", ep * num_experts * 4).view(torch.int32)[: ep * num_experts]
    cbuf[r * num_experts : (r + 1) * num_experts].copy_(splits)
    ctx.symm.all_gather_inplace(cbuf, num_experts, CH_EP_COUNTS, 1)
    counts = cbuf.view(ep, num_experts).to("cpu", non_blocking=False).to(torch.int64)  # the one host sync
    input_splits, output_splits, total_recv, cumsum_host, fwd, bwd = plan_chunks(counts, r, hidden * 2)
    cumsum_local = cumsum_host.to(torch.int32).to(dev, non_blocking=True)
    st = ctx.stats
    st["calls"] += 1
    st["sum_recv"] += total_recv
    st["max_recv"] = max(st["max_recv"], total_recv)
    st["min_recv"] = total_recv if st["min_recv"] is None else min(st["min_recv"], total_recv)
    return EPPlan(ctx, T, K, hidden, sidx, counts, input_splits, output_splits, total_recv, cumsum_local,
                  fwd.to(dev, non_blocking=True), fwd.shape[0], bwd.to(dev, non_blocking=True), bwd.shape[0])


def plan_chunks(counts: torch.Tensor, r: int, row: int):
    """Host geometry of one EP exchange from the gathered counts matrix ``counts[s, e]`` (int64, CPU) = rows rank ``s``
    sends to expert ``e``; ``row`` = bytes per token row. Vectorised (no per-expert Python loops: Qwen3-30B-A3B calls this
    48 x 3 times per step). Returns (input_splits, output_splits, total_recv, inclusive cumsum of rows per local expert,
    forward chunk list, return chunk list); chunk lists are int64 ``[n, 4]`` = {src_off, dst_off, bytes, peer} as the C
    struct ``ChunkDesc`` (csrc/p2p.cu), zero-sized blocks included (the kernel skips them), in
    expert-major / source-minor order — the order ``sort_chunks_by_idxs`` gives the reference (moe_utils.py:81-99)."""
    ep, E = counts.shape
    el = E // ep
    excl = torch.cumsum(counts, dim=1) - counts  # [EP, E]: row offset of expert e inside rank s's send order
    input_splits = counts[r].view(ep, el).sum(dim=1).tolist()
    mine = counts[:, r * el : (r + 1) * el]  # [EP, El]: what every source sends to my experts
    output_splits = mine.sum(dim=1).tolist()
    # forward: my receive buffer is ordered (local expert, source); block (le, s) comes from rank s at excl[s, e]
    n_fwd = mine.t().contiguous().view(-1)  # [El * EP] rows of block (le, s)
    dst_fwd = torch.cumsum(n_fwd, 0) - n_fwd
    total_recv = int(n_fwd.sum())
    src_fwd = excl[:, r * el : (r + 1) * el].t().contiguous().view(-1)
    peer_fwd = torch.arange(ep, dtype=torch.int64).repeat(el)
    fwd = torch.stack([src_fwd * row, dst_fwd * row, n_fwd * row, peer_fwd], dim=1)
    cumsum_host = torch.cumsum(mine.sum(dim=0), dim=0)
    # return path: for expert e = p*el + le (owner p) my block sits in p's buffer at
    #   rows(sum_{le' < le} sum_s C[s, p*el + le']) + sum_{s < r} C[s, e]   and goes back to row excl[r, e] of mine
    per_expert = counts.sum(dim=0).view(ep, el)  # [p, le]
    base = (torch.cumsum(per_expert, dim=1) - per_expert).view(-1)  # rows before expert e inside its owner's buffer
    before_me = counts[:r].sum(dim=0)  # [E]
    bwd = torch.stack([(base + before_me) * row, excl[r] * row, counts[r] * row,
                       torch.arange(ep, dtype=torch.int64).repeat_interleave(el)], dim=1)
    return input_splits, output_splits, total_recv, cumsum_host, fwd.contiguous(), bwd.contiguous()


def _staged(c: "EPContext", tag: str, rows: int, H: int):
    """(staging buffer, its [rows, H] bf16 view). The buffer itself is what a collective publishes: a rank that received no
    token at all (``rows == 0``: routing collapse, tiny batches) still takes part in the exchange, and an empty view has no
    address to derive the offset from."""
    buf = c.staging(tag, max(rows, 1) * H * 2)
    return buf, buf.view(BF)[: rows * H].view(-1, H)


def _pull(ctx: EPContext, channel: int, src_buf: torch.Tensor, chunks: torch.Tensor, n: int, out: torch.Tensor) -> None:
    """``src_buf`` = this rank's staging buffer (never empty); ``out`` may be empty."""
    lib = _lib.load()
    with torch.cuda.device(out.device), prof.span("ep_pull", out.numel() * out.element_size() * (ctx.ep_size - 1) / ctx.ep_size):
        check(lib.vb200_chunk_pull(ctx.symm.comm, channel, ctx.symm.offset_of(src_buf), chunks.data_ptr(), n,
                                   out.data_ptr(), ctx.num_ctas, stream_ptr()), "vb200_chunk_pull")


def _scatter_into(x: torch.Tensor, sidx: torch.Tensor, out: torch.Tensor, w: torch.Tensor | None = None) -> None:
    """out[sidx[t,k]] = x[t] (optionally scaled by w[t,k], rounded to bf16)."""
    T, K = sidx.shape
    lib = _lib.load()
    with torch.cuda.device(x.device):
        check(lib.vb200_moe_scatter(x.data_ptr(), sidx.data_ptr(), out.data_ptr(), w.data_ptr() if w is not None else None,
                                    None, T, K, x.shape[-1], stream_ptr()), "vb200_moe_scatter")


def routing_weight_grad(g: torch.Tensor, rows: torch.Tensor, sidx: torch.Tensor) -> torch.Tensor:
    """fp32 ``[T, K]``: ``<g[t], rows[sidx[t,k]]>`` — d(loss)/d(routing weight) of the weighted combine. One kernel (a warp
    per token keeps ``g[t]`` in registers for its K rows) instead of a gathered ``[T, K, H]`` copy, two fp32 casts and an
    einsum; hidden sizes that are not a multiple of 256 (toy models) keep the torch expression."""
    T, K = sidx.shape
    H = g.shape[-1]
    if H % 256 or H > 8192 or (H // 256 > 8 and H // 256 not in (16, 20, 32)):
        return torch.einsum("th,tkh->tk", g.float(), rows[sidx.flatten().long()].view(T, K, H).float())
    out = torch.empty(T, K, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        check(_lib.load().vb200_moe_weight_grad(g.data_ptr(), rows.data_ptr(), sidx.data_ptr(), out.data_ptr(), T, K, H,
                                                stream_ptr()), "vb200_moe_weight_grad")
    return out


class _EPDispatch(torch.autograd.Function):
    """token_pre_all2all (moe_layer.py:72-99): local permute -> exchange -> group by local expert."""

    @staticmethod
    def forward(ctx, hs, plan: EPPlan):
        c = plan.ctx
        sbuf, send = _staged(c, "dispatch_send", plan.T * plan.K, plan.H)
        _scatter_into(hs.contiguous(), plan.sidx, send)
        recv = torch.empty(max(plan.total_recv, 1), plan.H, dtype=BF, device=hs.device)[: plan.total_recv]
        _pull(c, CH_EP_DISPATCH, sbuf, plan.fwd_chunks, plan.n_fwd, recv)
        ctx.plan = plan
        return recv

    @staticmethod
    def backward(ctx, g):
        plan: EPPlan = ctx.plan
        c = plan.ctx
        # gradients of received rows go back to their sources, then sum over the top-k slots
        rbuf, ret = _staged(c, "dispatch_grad", plan.total_recv, plan.H)
        ret.copy_(g)
        back = torch.empty(plan.T * plan.K, plan.H, dtype=BF, device=g.device)
        _pull(c, CH_EP_COMBINE, rbuf, plan.bwd_chunks, plan.n_bwd, back)
        return _gather_raw(back, plan.sidx), None


class _EPCombine(torch.autograd.Function):
    """tokens_post_all2all (moe_layer.py:102-137): exchange back -> weight -> fp32 sum over the top-k."""

    @staticmethod
    def forward(ctx, expert_out, weights, plan: EPPlan):
        c = plan.ctx
        rbuf, ret = _staged(c, "combine_ret", plan.total_recv, plan.H)
        ret.copy_(expert_out)
        back = torch.empty(plan.T * plan.K, plan.H, dtype=BF, device=expert_out.device)
        _pull(c, CH_EP_COMBINE, rbuf, plan.bwd_chunks, plan.n_bwd, back)
        w = weights.to(BF).contiguous()
        out = _gather_raw(back, plan.sidx, w)
        ctx.plan = plan
        ctx.save_for_backward(back, w)
        ctx.w_dtype = weights.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        plan: EPPlan = ctx.plan
        c = plan.ctx
        back, w = ctx.saved_tensors
        g = g.contiguous()
        # d/d(expert_out): scatter w[t,k] * g[t] into the permuted order, owners pull their rows
        sbuf, send = _staged(c, "combine_grad", plan.T * plan.K, plan.H)
        _scatter_into(g, plan.sidx, send, w)
        grad_expert = torch.empty(max(plan.total_recv, 1), plan.H, dtype=BF, device=g.device)[: plan.total_recv]
        _pull(c, CH_EP_DISPATCH, sbuf, plan.fwd_chunks, plan.n_fwd, grad_expert)
        # d/d(weights)[t,k] = <g[t], back[sidx[t,k]]>
        grad_w = routing_weight_grad(g, back, plan.sidx).to(ctx.w_dtype)
        return grad_expert, grad_w, None


class _EPExperts(torch.autograd.Function):
    """EPMergedFc1GroupGemm (moe_layer.py:307-441): fc1 -> silu*up -> fc2 on the local experts, recompute in bwd."""

    @staticmethod
    def forward(ctx, tokens, cumsum, w1, w2):
        inter = w1.shape[1] // 2
        fc1 = group_gemm_same_nk(tokens, w1, cumsum, transpose_b=True)
        with torch.no_grad():
            act = F.silu_mul(fc1[:, :inter], fc1[:, inter:])
        out = group_gemm_same_nk(act, w2, cumsum, transpose_b=True)
        ctx.save_for_backward(tokens, cumsum, w1, w2, fc1)
        return out

    @staticmethod
    def backward(ctx, g):
        tokens, cumsum, w1, w2, fc1 = ctx.saved_tensors
        inter = w1.shape[1] // 2
        g = g.contiguous()
        with torch.no_grad():
            act = F.silu_mul(fc1[:, :inter], fc1[:, inter:])  # recompute (moe_layer.py:389-391)
        grad_act = group_gemm_same_nk(g, w2, cumsum, transpose_b=False)
        grad_w2 = None
        if ctx.needs_input_grad[3]:
            grad_w2 = torch.empty_like(w2)
            group_gemm_same_mn(g, act, grad_w2, cumsum)
        grad_fc1 = torch.empty_like(fc1)
        lib = _lib.load()
        with torch.cuda.device(g.device):
            check(lib.vb200_swiglu_bwd(grad_act.data_ptr(), fc1.data_ptr(), fc1.data_ptr() + inter * 2, grad_fc1.data_ptr(),
                                       grad_fc1.data_ptr() + inter * 2, fc1.shape[0], inter, 2 * inter, inter, 2 * inter,
                                       stream_ptr()), "vb200_swiglu_bwd")
        grad_tokens = group_gemm_same_nk(grad_fc1, w1, cumsum, transpose_b=False)
        grad_w1 = None
        if ctx.needs_input_grad[2]:
            grad_w1 = torch.empty_like(w1)
            group_gemm_same_mn(grad_fc1, tokens, grad_w1, cumsum)
        return grad_tokens, None, grad_w1, grad_w2


def ep_dispatch(ctx: EPContext, hidden_states: torch.Tensor, selected_experts: torch.Tensor, num_experts: int):
    """Returns (tokens [sum_recv, H] grouped by local expert, plan)."""
    hs = hidden_states.reshape(-1, hidden_states.shape[-1])
    plan = make_plan(ctx, selected_experts, num_experts, hs.shape[-1])
    return _EPDispatch.apply(hs, plan), plan


def ep_combine(expert_outputs: torch.Tensor, routing_weights: torch.Tensor, plan: EPPlan) -> torch.Tensor:
    return _EPCombine.apply(expert_outputs, routing_weights, plan)


def ep_fused_moe_forward(ctx: EPContext, num_experts: int, routing_weights, selected_experts, hidden_states,
                         fc1_1_2_weight_local, fc2_weight_local):
    """EP branch of the fused MoE forward; weights are this rank's ``[E/EP, ...]`` slices (ParallelPlan Shard(0))."""
    shape = hidden_states.shape
    tokens, plan = ep_dispatch(ctx, hidden_states, selected_experts, num_experts)
    out = _EPExperts.apply(tokens, plan.cumsum_local, fc1_1_2_weight_local, fc2_weight_local)
    return ep_combine(out, routing_weights, plan).reshape(shape)
