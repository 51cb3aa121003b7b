"""Oracle for the causal-LM loss (cross-entropy over the vocabulary).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.

Restates, on torch CPU tensors:
* ``fixed_cross_entropy`` as called by ``eager_cross_entropy``
  (veomni/ops/kernels/cross_entropy/eager.py:23-38; transformers/loss/loss_utils.py): mean over the
  non-ignored rows, or sum / num_items_in_batch;
* the label shift of ``ForCausalLMLoss`` (veomni/ops/kernels/cross_entropy/__init__.py:180-190);
* the fused-linear form (liger ``LigerFusedLinearCrossEntropyLoss`` bound by
  veomni/ops/kernels/cross_entropy/liger.py:21-57 — third-party, absent from /root/reference; its published
  algorithm: per row chunk, logits = h @ W^T in the compute dtype, softmax statistics in fp32, the gradient
  (softmax - onehot) / n_valid written over the logits in the compute dtype and immediately contracted into
  dH and dW). Pinned against the reference's eager path on the same inputs in tests/golden/make_golden.py.
"""

from __future__ import annotations

import torch


def shift_labels(labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """ForCausalLMLoss (__init__.py:184-186): tokens < n predict n."""
    padded = torch.nn.functional.pad(labels, (0, 1), value=ignore_index)
    return padded[..., 1:].contiguous()


def cross_entropy_rows(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100):
    """Per-row loss logsumexp(x) - x[label] in fp32 (0 for ignored rows) and the row logsumexp."""
    x = logits.to(torch.float32)
    m = x.max(dim=-1, keepdim=True).values
    lse = (m + (x - m).exp().sum(dim=-1, keepdim=True).log()).squeeze(-1)
    valid = labels != ignore_index
    picked = x.gather(1, labels.clamp_min(0)[:, None]).squeeze(1)
    return torch.where(valid, lse - picked, torch.zeros_like(lse)), lse


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, num_items_in_batch=None, ignore_index: int = -100):
    """fixed_cross_entropy: reduction "mean" over valid rows, or "sum" / num_items_in_batch."""
    rows, _ = cross_entropy_rows(logits, labels, ignore_index)
    if num_items_in_batch is None:
        return rows.sum() / (labels != ignore_index).sum().to(torch.float32)
    return rows.sum() / float(num_items_in_batch)


def cross_entropy_grad(logits: torch.Tensor, labels: torch.Tensor, scale: float, ignore_index: int = -100):
    """d(scale * sum_rows loss_row) / d logits, in fp32: (softmax - onehot) * scale, 0 for ignored rows."""
    x = logits.to(torch.float32)
    p = torch.softmax(x, dim=-1)
    valid = labels != ignore_index
    p[torch.arange(x.size(0))[valid], labels[valid]] -= 1.0
    p[~valid] = 0.0
    return p * scale


def fused_linear_cross_entropy(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                               num_items_in_batch=None, ignore_index: int = -100, chunk_size: int = 1024):
    """Chunked lm_head + cross-entropy; returns (loss, d hidden, d weight) with the gradients rounded to the compute
    dtype exactly where the fused kernels round them (the logits chunk and its in-place gradient)."""
    T = hidden.size(0)
    n_valid = (labels != ignore_index).sum().to(torch.float32)
    scale = (1.0 / n_valid).item() if num_items_in_batch is None else 1.0 / float(num_items_in_batch)
    total = torch.zeros((), dtype=torch.float32)
    dh = torch.zeros_like(hidden)
    dw = torch.zeros_like(weight)
    for r0 in range(0, T, chunk_size):
        r1 = min(T, r0 + chunk_size)
        logits = hidden[r0:r1] @ weight.t()  # compute dtype
        rows, _ = cross_entropy_rows(logits, labels[r0:r1], ignore_index)
        total = total + rows.sum()
        g = cross_entropy_grad(logits, labels[r0:r1], scale, ignore_index).to(hidden.dtype)
        dh[r0:r1] = g @ weight
        dw += g.t() @ hidden[r0:r1]
    return total * scale, dh, dw


def reduce_sequence_parallel_loss(losses: list[float], n_valid: list[int], upstream: float = 1.0):
    """``reduce_sequence_parallel_loss`` / ``ReduceLoss`` (veomni/distributed/sequence_parallel/loss.py:27-64) over the
    ranks of one SP group: forward sum_r(loss_r * n_r) / max(sum_r n_r, 1) with ranks that hold no valid token
    contributing 0; backward d/d loss_r = world * n_r * upstream / max(sum n, 1). Returns (reduced, [grad per rank])."""
    world = len(losses)
    total = sum((l if n > 0 else 0.0) * n for l, n in zip(losses, n_valid))
    denom = max(sum(n_valid), 1)
    return total / denom, [world * n * upstream / denom for n in n_valid]
