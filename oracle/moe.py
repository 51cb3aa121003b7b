"""Oracle for MoE routing, GroupGEMM and expert-parallel dispatch / combine.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.

Integer routing (histogram, scatter index, split sizes, permutation mappings) is restated
exactly; parity for it is bit-exact.  The Triton GroupGEMM / scatter / gather kernels cannot run
without CUDA, so their arithmetic is restated from the kernel sources and docstrings (bf16
operands, fp32 accumulation, one rounding of the result) — the reference tests that pin them are
tests/ops/test_fused_moe_split_vs_merged.py:145-161,362-486.

Multi-rank functions take *lists indexed by EP rank* and perform the all-to-all by slicing, which
is the definition of ``dist.all_to_all_single`` with split sizes
(veomni/distributed/moe/comm.py:20-54).
"""

from __future__ import annotations

import torch

F = torch.nn.functional


# ---------------------------------------------------------------------------------------------
# single-GPU routing (veomni/ops/kernels/moe/group_gemm.py, _kernels/kernel/moe.py)
# ---------------------------------------------------------------------------------------------
def expert_histogram(expert_index: torch.Tensor, num_experts: int) -> torch.Tensor:
    """Tokens per expert, int32 [E].  Reference: _kernels/kernel/moe.py:53-82 (bin width 1)."""
    return torch.bincount(expert_index.flatten().to(torch.int64), minlength=num_experts)[:num_experts].to(torch.int32)


def scatter_index(expert_index: torch.Tensor) -> torch.Tensor:
    """Row of each (token, k) slot in the expert-sorted activation, int32 [T, K].

    Reference: veomni/ops/kernels/moe/group_gemm.py:44 and :287 —
    ``expert_index.flatten().argsort(stable=True).argsort().int().view(expert_index.shape)``.
    """
    return expert_index.flatten().argsort(stable=True).argsort().int().view(expert_index.shape)


def moe_scatter(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``O[index[m, k]] = X[m]``.  Reference: _kernels/kernel/moe.py:270-300 (docstring :270-281)."""
    T, K = index.shape
    out = torch.empty(T * K, x.shape[1], dtype=x.dtype)
    out[index.flatten().long()] = x.repeat_interleave(K, dim=0)
    return out


def moe_gather(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``Y[m] = sum_k X[index[m, k]]`` accumulated in fp32 in k order, rounded once.

    Reference: _kernels/kernel/moe.py:104-127 (``y = tl.zeros(fp32)``; ``y += x`` over TOPK)."""
    T, K = index.shape
    acc = torch.zeros(T, x.shape[1], dtype=torch.float32)
    for k in range(K):
        acc += x[index[:, k].long()].float()
    return acc.to(x.dtype)


def _group_bounds(cumsum: torch.Tensor):
    ends = [int(v) for v in cumsum.tolist()]
    starts = [0] + ends[:-1]
    return list(zip(starts, ends))


def group_gemm_same_nk(a: torch.Tensor, b: torch.Tensor, cumsum_M: torch.Tensor, transpose_b: bool) -> torch.Tensor:
    """Ragged-M GroupGEMM: rows ``cumsum_M[g-1]:cumsum_M[g]`` of A times expert g of B.

    Reference: _kernels/kernel/group_gemm.py:157-234 (kernel :54-154): ``transpose_b=True`` means
    ``b`` is ``[G, N, K]`` and C = A @ B[g]^T; ``False`` means ``[G, K, N]`` and C = A @ B[g].
    fp32 accumulation, output in a.dtype.  Rows past ``cumsum_M[-1]`` are left unspecified by the
    reference (never written); the oracle zero-fills them.
    """
    N = b.shape[1] if transpose_b else b.shape[2]
    c = torch.zeros(a.shape[0], N, dtype=a.dtype)
    for g, (s, e) in enumerate(_group_bounds(cumsum_M)):
        if e > s:
            w = b[g].float()
            c[s:e] = (a[s:e].float() @ (w.t() if transpose_b else w)).to(a.dtype)
    return c


def group_gemm_same_mn(a: torch.Tensor, b: torch.Tensor, cumsum_K: torch.Tensor) -> torch.Tensor:
    """Ragged-K GroupGEMM (wgrad): ``C[g] = A[rows g]^T @ B[rows g]``, zero when the group is empty.

    Reference: _kernels/kernel/group_gemm.py:357-397 with ``transpose_a=True, transpose_b=False``
    (kernel :241-354, zero-fill for k == 0 at :323-337).  a: [sumK, M], b: [sumK, N] -> [G, M, N].
    """
    G = cumsum_K.numel()
    c = torch.zeros(G, a.shape[1], b.shape[1], dtype=a.dtype)
    for g, (s, e) in enumerate(_group_bounds(cumsum_K)):
        if e > s:
            c[g] = (a[s:e].float().t() @ b[s:e].float()).to(a.dtype)
    return c


def fused_moe_forward(num_experts, routing_weights, selected_experts, hidden_states, fc1_1_2_weight, fc2_weight):
    """Non-EP fused MoE forward (merged fc1).

    Reference: MergedFc1TritonFusedMoeExpertFunction.forward,
    veomni/ops/kernels/moe/group_gemm.py:277-345 — note the routing weight is applied to the fc1
    activation *before* fc2 (:306-311).  Returns (output, intermediates dict).
    """
    splits = expert_histogram(selected_experts, num_experts)
    sidx = scatter_index(selected_experts)
    scatter_output = moe_scatter(hidden_states, sidx)
    cumsum_t = torch.cumsum(splits, dim=0)
    fc1_output = group_gemm_same_nk(scatter_output, fc1_1_2_weight, cumsum_t, transpose_b=True)
    fc1_1_output, fc1_2_output = fc1_output.chunk(2, dim=-1)
    fc1_activation = F.silu(fc1_1_output) * fc1_2_output
    reshaped_gate_weight = routing_weights.reshape(-1, 1)
    scattered_gate_weight = torch.empty_like(reshaped_gate_weight)
    scattered_gate_weight[sidx.flatten().long()] = reshaped_gate_weight
    fc1_weighted_output = fc1_activation * scattered_gate_weight
    fc2_output = group_gemm_same_nk(fc1_weighted_output, fc2_weight, cumsum_t, transpose_b=True)
    output = moe_gather(fc2_output, sidx).reshape(hidden_states.shape)
    return output, {
        "splits": splits, "scatter_index": sidx, "scatter_output": scatter_output, "cumsum": cumsum_t,
        "fc1_output": fc1_output, "fc2_output": fc2_output,
    }


def eager_moe_forward(num_experts, routing_weights, selected_experts, hidden_states, fc1_1_2_weight, fc2_weight):
    """HF eager expert loop (the reference tests' ground truth).

    Reference: Qwen3MoeExperts.forward, veomni/models/transformers/qwen3_moe/generated/
    patched_modeling_qwen3_moe_gpu.py:276-300 and tests/ops/test_fused_moe_split_vs_merged.py:21-44.
    """
    out = torch.zeros_like(hidden_states)
    expert_mask = F.one_hot(selected_experts, num_classes=num_experts).permute(2, 1, 0)
    expert_hit = torch.greater(expert_mask.sum(dim=(-1, -2)), 0).nonzero()
    for expert_idx in expert_hit:
        idx = int(expert_idx[0])
        top_k_pos, token_idx = torch.where(expert_mask[idx])
        cur = hidden_states[token_idx]
        gate, up = F.linear(cur, fc1_1_2_weight[idx]).chunk(2, dim=-1)
        y = F.linear(F.silu(gate) * up, fc2_weight[idx])
        y = y * routing_weights[token_idx, top_k_pos, None]
        out.index_add_(0, token_idx, y.to(out.dtype))
    return out


# ---------------------------------------------------------------------------------------------
# expert parallel (veomni/distributed/moe/moe_layer.py, moe_utils.py, comm.py)
# ---------------------------------------------------------------------------------------------
def expert_mask_of(selected_experts: torch.Tensor, num_experts: int) -> torch.Tensor:
    """``one_hot(selected_experts).permute(2, 1, 0)`` -> [E, K, T] (group_gemm.py:474)."""
    return F.one_hot(selected_experts, num_classes=num_experts).permute(2, 1, 0)


def preprocess(expert_masks: list[torch.Tensor], num_experts: int):
    """Split sizes of the EP exchange for every rank.

    Reference: veomni/distributed/moe/moe_layer.py:30-69.  Returns per-rank lists of
    (input_splits [EP], output_splits [EP], num_global_tokens_per_local_expert [EP, E/EP],
    num_global_sum_tokens_per_local_expert [E/EP]).
    """
    ep = len(expert_masks)
    nle = num_experts // ep
    local = [m.sum(dim=(1, 2)) for m in expert_masks]  # [E] per rank
    gathered = torch.stack(local, dim=0)  # all_gather_into_tensor -> [EP, E]
    res = []
    for r in range(ep):
        input_splits = local[r].reshape(ep, nle).sum(dim=1).tolist()
        per_local = gathered[:, r * nle : (r + 1) * nle].contiguous()
        output_splits = per_local.sum(dim=1).tolist()
        res.append((input_splits, output_splits, per_local.view(-1, nle), per_local.sum(dim=0)))
    return res


def permute(tokens: torch.Tensor, routing_map: torch.Tensor):
    """Expert-major local permutation.  Reference: veomni/distributed/moe/moe_utils.py:19-41."""
    num_tokens = tokens.shape[0]
    num_experts = routing_map.shape[0]
    rm = routing_map.bool()
    token_indices = torch.arange(num_tokens).unsqueeze(0).expand(num_experts, -1)
    sorted_indices = token_indices.masked_select(rm)
    return tokens.index_select(0, sorted_indices), sorted_indices


def generate_weights_idx(routing_weights, selected_experts, num_experts):
    """Reference: moe_utils.py:75-92."""
    num_tokens = routing_weights.shape[0]
    w = torch.zeros((num_tokens, num_experts), dtype=routing_weights.dtype)
    w.scatter_add_(1, selected_experts, routing_weights)
    return w


def unpermute(tokens, routing_weights_dense, hidden_states_shape, permutation_mapping, routing_map):
    """Weight and scatter-add back in fp32.  Reference: moe_utils.py:44-72."""
    tokens_weight = routing_weights_dense.T.contiguous().masked_select(routing_map.bool())
    tokens = tokens * tokens_weight.unsqueeze(-1)
    hidden_dim = hidden_states_shape[-1]
    out = torch.zeros(hidden_states_shape, dtype=torch.float32)
    out.scatter_add_(0, permutation_mapping.unsqueeze(1).expand(-1, hidden_dim), tokens.float())
    return out.to(tokens.dtype)


def sort_chunks_by_idxs(x, split_sizes, sorted_idxs):
    """Reference: moe_utils.py:95-99."""
    chunks = torch.split(x, [int(s) for s in split_sizes.tolist()], dim=0)
    return torch.cat([chunks[i] for i in sorted_idxs], dim=0)


def all_to_all(inputs: list[torch.Tensor], input_splits: list[list[int]]) -> list[torch.Tensor]:
    """``all_to_all_single`` with split sizes over a list of ranks (comm.py:20-54)."""
    ep = len(inputs)
    chunks = [torch.split(inputs[s], input_splits[s], dim=0) for s in range(ep)]
    return [torch.cat([chunks[s][r] for s in range(ep)], dim=0) for r in range(ep)]


def ep_dispatch(hidden_states: list[torch.Tensor], selected_experts: list[torch.Tensor], num_experts: int):
    """token_pre_all2all on every rank (moe_layer.py:72-99). Returns per-rank dict of results."""
    ep = len(hidden_states)
    nle = num_experts // ep
    masks = [expert_mask_of(se, num_experts) for se in selected_experts]
    pre = preprocess(masks, num_experts)
    local_perm, mappings, routing_maps = [], [], []
    for r in range(ep):
        routing_map = masks[r].sum(dim=1)
        p, m = permute(hidden_states[r].reshape(-1, hidden_states[r].shape[-1]), routing_map)
        local_perm.append(p)
        mappings.append(m)
        routing_maps.append(routing_map)
    recv = all_to_all(local_perm, [pre[r][0] for r in range(ep)])
    permute_order = torch.arange(num_experts).reshape(-1, nle).T.ravel().tolist()
    out = []
    for r in range(ep):
        tokens = sort_chunks_by_idxs(recv[r], pre[r][2].ravel(), permute_order)
        out.append({
            "tokens": tokens, "input_splits": pre[r][0], "output_splits": pre[r][1],
            "num_global_tokens_per_local_expert": pre[r][2],
            "num_global_sum_tokens_per_local_expert": pre[r][3],
            "routing_map": routing_maps[r], "permutation_mapping": mappings[r],
            "cumsum": torch.cumsum(pre[r][3], dim=0),
        })
    return out


def ep_combine(expert_outputs: list[torch.Tensor], dispatched: list[dict], routing_weights: list[torch.Tensor],
               selected_experts: list[torch.Tensor], num_experts: int, shapes: list[torch.Size]):
    """tokens_post_all2all on every rank (moe_layer.py:102-137)."""
    ep = len(expert_outputs)
    nle = num_experts // ep
    unpermute_order = torch.arange(num_experts).reshape(nle, -1).T.ravel().tolist()
    sorted_out = [
        sort_chunks_by_idxs(expert_outputs[r], dispatched[r]["num_global_tokens_per_local_expert"].T.ravel(),
                            unpermute_order)
        for r in range(ep)
    ]
    back = all_to_all(sorted_out, [dispatched[r]["output_splits"] for r in range(ep)])
    res = []
    for r in range(ep):
        w = generate_weights_idx(routing_weights[r], selected_experts[r], num_experts)
        res.append(unpermute(back[r], w, shapes[r], dispatched[r]["permutation_mapping"], dispatched[r]["routing_map"]))
    return res


def ep_expert_mlp(tokens, cumsum, fc1_1_2_weight, fc2_weight):
    """EPMergedFc1GroupGemm.forward (moe_layer.py:314-372): fc1 -> silu*up -> fc2, no routing weight."""
    fc1 = group_gemm_same_nk(tokens, fc1_1_2_weight, cumsum, transpose_b=True)
    g, u = fc1.chunk(2, dim=-1)
    return group_gemm_same_nk(F.silu(g) * u, fc2_weight, cumsum, transpose_b=True)


def ep_moe_forward(hidden_states, routing_weights, selected_experts, num_experts, fc1_1_2_weight, fc2_weight):
    """Whole EP branch of group_gemm_fused_moe_forward (group_gemm.py:468-524) for all ranks.

    ``fc1_1_2_weight`` / ``fc2_weight`` are the *global* ``[E, ...]`` weights; rank r owns the
    contiguous block ``[r*E/EP, (r+1)*E/EP)`` (ParallelPlan Shard(0),
    veomni/distributed/parallel_plan.py:53-101).  The routing weight is applied after the return
    all-to-all, in ``unpermute``.
    """
    ep = len(hidden_states)
    nle = num_experts // ep
    disp = ep_dispatch(hidden_states, selected_experts, num_experts)
    outs = [
        ep_expert_mlp(disp[r]["tokens"], disp[r]["cumsum"], fc1_1_2_weight[r * nle : (r + 1) * nle],
                      fc2_weight[r * nle : (r + 1) * nle])
        for r in range(ep)
    ]
    shapes = [h.reshape(-1, h.shape[-1]).shape for h in hidden_states]
    return ep_combine(outs, disp, routing_weights, selected_experts, num_experts, shapes), disp
