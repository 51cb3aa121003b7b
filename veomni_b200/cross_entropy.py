"""Cross-entropy over the vocabulary: the inner kernel of VeOmni's loss wrappers.

Mirrors veomni/ops/kernels/cross_entropy/:
* :func:`b200_cross_entropy` has the ``cross_entropy_fn`` contract of ``eager_cross_entropy`` (eager.py:23-38) and
  ``fused_liger_kernel_cross_entropy`` (liger.py:24-57): ``(logits, labels, vocab_size, num_items_in_batch,
  ignore_index, shift_labels, hidden_states=, weights=) -> (loss, logits)``. With ``logits`` it is the eager
  arithmetic (``fixed_cross_entropy``: mean over non-ignored rows, or sum / num_items_in_batch); with
  ``hidden_states`` + ``weights`` it is the fused-linear form: the ``lm_head`` projection runs chunk by chunk on cuBLAS,
  the CUDA kernel turns each logits chunk into its gradient in place, and the full ``[T, V]`` logits (2.5 GB in
  fp32 at T=4096, V=151936) never exist.
* :func:`ForCausalLMLoss` is the outer policy (__init__.py:89-221): label shift unless SP, flatten, SP loss reduce.

No host synchronisation: the valid-token count stays on the device and scales loss and gradient there.
"""

from __future__ import annotations

from typing import Any, Callable

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import VB200Error, check, stream_ptr

_DT = {torch.bfloat16: 0, torch.float32: 1}


def _check(logits: torch.Tensor, labels: torch.Tensor) -> None:
    if not logits.is_cuda or not labels.is_cuda:
        raise VB200Error("cross_entropy runs on CUDA tensors only (no CPU fallback)")
    if logits.dtype not in _DT:
        raise VB200Error(f"cross_entropy: logits must be bf16 or fp32, got {logits.dtype}")
    if logits.dim() != 2 or logits.stride(1) != 1:
        raise VB200Error("cross_entropy: logits must be [rows, vocab] with contiguous rows")
    if labels.dtype != torch.int64 or labels.dim() != 1 or labels.numel() != logits.size(0):
        raise VB200Error("cross_entropy: labels must be int64 [rows]")


def _ptr(t: torch.Tensor | None) -> int | None:
    return t.data_ptr() if t is not None else None


def _launch(logits, labels, ignore_index, loss_rows, lse, lse_given, grad, scale, scale_dev, upstream) -> None:
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        check(lib.vb200_cross_entropy(
            logits.data_ptr(), _DT[logits.dtype], logits.size(0), logits.size(1), logits.stride(0), labels.data_ptr(),
            int(ignore_index), _ptr(loss_rows), _ptr(lse), int(lse_given), _ptr(grad),
            grad.stride(0) if grad is not None else 0, float(scale), _ptr(scale_dev), _ptr(upstream), stream_ptr(),
        ), "vb200_cross_entropy")


def valid_label_recip(labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """Device tensor ``[1/count, count]`` of labels != ignore_index. With no valid label the factor is 0, so the loss and
    its gradient are exactly 0 for that batch (csrc/cross_entropy.cu ``ce_valid_recip_kernel``) — NOT the NaN that
    ``F.cross_entropy(reduction="mean")`` returns for an all-ignored batch: a fully masked micro-batch does not poison
    the step. Callers that want the eager NaN can test ``out[1] == 0`` themselves."""
    out = torch.empty(2, dtype=torch.float32, device=labels.device)
    lab = labels.reshape(-1).contiguous()
    if not lab.is_cuda or lab.dtype != torch.int64:
        raise VB200Error("valid_label_recip: labels must be a CUDA int64 tensor")
    with torch.cuda.device(lab.device):
        check(_lib.load().vb200_count_valid_labels(lab.data_ptr(), lab.numel(), int(ignore_index), out.data_ptr(),
                                                   stream_ptr()), "vb200_count_valid_labels")
    return out


def _scales(labels: torch.Tensor, num_items_in_batch, ignore_index: int) -> tuple[float, torch.Tensor | None]:
    """(host factor, device factor) of the reduction: mean over valid rows, or sum / num_items_in_batch
    (transformers fixed_cross_entropy)."""
    if num_items_in_batch is None:
        return 1.0, valid_label_recip(labels, ignore_index)[:1]
    if torch.is_tensor(num_items_in_batch):
        return 1.0, (1.0 / num_items_in_batch.to(device=labels.device, dtype=torch.float32)).reshape(1)
    return 1.0 / float(num_items_in_batch), None


class _CrossEntropy(torch.autograd.Function):
    """loss = reduce(logsumexp(x) - x[label]); backward re-reads the logits with the saved logsumexp."""

    @staticmethod
    def forward(ctx: Any, logits: torch.Tensor, labels: torch.Tensor, ignore_index: int, scale: float,
                scale_dev: torch.Tensor | None):
        _check(logits, labels)
        rows = logits.size(0)
        loss_rows = torch.empty(rows, dtype=torch.float32, device=logits.device)
        lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
        _launch(logits, labels, ignore_index, loss_rows, lse, 0, None, 1.0, None, None)
        loss = loss_rows.sum() * scale
        if scale_dev is not None:
            loss = loss * scale_dev[0]
        ctx.save_for_backward(logits, labels, lse, scale_dev)
        ctx.ignore_index, ctx.scale = ignore_index, scale
        return loss

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):
        logits, labels, lse, scale_dev = ctx.saved_tensors
        grad = torch.empty_like(logits)
        up = g.detach().to(torch.float32).reshape(1).contiguous()
        _launch(logits, labels, ctx.ignore_index, None, lse, 1, grad, ctx.scale, scale_dev, up)
        return grad, None, None, None, None


class _FusedLinearCrossEntropy(torch.autograd.Function):
    """lm_head projection + cross-entropy, chunked over rows; the gradients w.r.t. hidden states and weights are
    produced during the forward pass (as the liger kernel the reference binds does) and scaled by the incoming
    gradient in backward."""

    @staticmethod
    def forward(ctx: Any, hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, ignore_index: int,
                scale: float, scale_dev: torch.Tensor | None, chunk_size: int):
        if hidden.dim() != 2 or weight.dim() != 2 or hidden.size(1) != weight.size(1):
            raise VB200Error("fused linear cross-entropy: hidden [T, H] and weight [V, H] expected")
        if hidden.dtype != weight.dtype:
            raise VB200Error("fused linear cross-entropy: hidden states and weights must share a dtype")
        T, V = hidden.size(0), weight.size(0)
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        loss_rows = torch.empty(T, dtype=torch.float32, device=hidden.device)
        grad_h = torch.empty_like(hidden) if need_h else None
        grad_w = torch.zeros_like(weight) if need_w else None
        wt = weight.t()
        for r0 in range(0, T, chunk_size):
            r1 = min(T, r0 + chunk_size)
            h_c = hidden[r0:r1]
            logits = torch.mm(h_c, wt)  # [chunk, V] in the compute dtype (cuBLAS)
            lab = labels[r0:r1]
            _check(logits, lab)
            grad = logits if (need_h or need_w) else None  # in place: the chunk becomes dLoss/dlogits
            _launch(logits, lab, ignore_index, loss_rows[r0:r1], None, 0, grad, scale, scale_dev, None)
            if need_h:
                torch.mm(logits, weight, out=grad_h[r0:r1])
            if need_w:
                grad_w.addmm_(logits.t(), h_c)
        loss = loss_rows.sum() * scale
        if scale_dev is not None:
            loss = loss * scale_dev[0]
        ctx.save_for_backward(grad_h, grad_w)
        return loss

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):
        grad_h, grad_w = ctx.saved_tensors
        gh = grad_h * g.to(grad_h.dtype) if grad_h is not None else None
        gw = grad_w * g.to(grad_w.dtype) if grad_w is not None else None
        return gh, gw, None, None, None, None, None


class _ReduceLoss(torch.autograd.Function):
    """Token-weighted mean of the per-rank losses over the SP group (sequence_parallel/loss.py:27-60):
    forward sum_r(loss_r * n_r) / max(sum_r n_r, 1), a rank without valid tokens contributing 0;
    backward world * n_local / max(n_global, 1) * g."""

    @staticmethod
    def forward(ctx: Any, loss: torch.Tensor, num_valid: torch.Tensor, group):
        loss = torch.where(num_valid > 0, loss, torch.zeros_like(loss))
        local_n = num_valid.detach().clone()
        total = loss * num_valid
        global_n = num_valid.detach().clone()
        torch.distributed.all_reduce(total, group=group)
        torch.distributed.all_reduce(global_n, group=group)
        ctx.save_for_backward(local_n, global_n)
        ctx.world = torch.distributed.get_world_size(group)
        return total / global_n.clamp_min(1)

    @staticmethod
    def backward(ctx: Any, g: torch.Tensor):
        local_n, global_n = ctx.saved_tensors
        return ctx.world * local_n * g / global_n.clamp(min=1), None, None


def b200_cross_entropy(
    logits: torch.Tensor | None = None,
    labels: torch.Tensor | None = None,
    vocab_size: int | None = None,
    num_items_in_batch: int | torch.Tensor | None = None,
    ignore_index: int = -100,
    shift_labels: torch.Tensor | None = None,
    **kwargs,
) -> tuple[torch.Tensor, torch.Tensor | None]:
    """``cross_entropy_fn`` for ForCausalLMLoss / ForSequenceClassificationLoss. Prefers the fused-linear form when the
    caller passes ``hidden_states`` and ``weights`` (then no logits are returned, as with the liger kernel)."""
    hidden_states = kwargs.pop("hidden_states", None)
    weights = kwargs.pop("weights", None)
    chunk_size = int(kwargs.pop("chunk_size", 1024))
    labels = labels.reshape(-1)
    scale, scale_dev = _scales(labels, num_items_in_batch, ignore_index)
    if hidden_states is not None and weights is not None:
        hidden_states = hidden_states.reshape(-1, hidden_states.size(-1))
        loss = _FusedLinearCrossEntropy.apply(hidden_states, weights, labels, ignore_index, scale, scale_dev, chunk_size)
        return loss, logits
    if logits is None:
        raise VB200Error("b200_cross_entropy needs logits, or hidden_states and weights")
    logits = logits.reshape(-1, vocab_size if vocab_size is not None else logits.size(-1))
    return _CrossEntropy.apply(logits, labels, ignore_index, scale, scale_dev), logits


def ForCausalLMLoss(
    logits: torch.Tensor | None = None,
    labels: torch.Tensor | None = None,
    vocab_size: int | None = None,
    num_items_in_batch: int | None = None,
    ignore_index: int = -100,
    shift_labels: torch.Tensor | None = None,
    *,
    cross_entropy_fn: Callable = b200_cross_entropy,
    sp_group=None,
    **kwargs,
):
    """Outer policy of the causal-LM loss (reference __init__.py:89-221, loss path only): shift labels unless the
    sequence is SP-sharded (then the data pipeline already shifted them), flatten, call the kernel, and reduce the
    loss over the SP group weighted by each rank's valid-token count (sequence_parallel/loss.py).
    Returns ``(loss, logits, None)`` like the reference wrapper."""
    hidden_states = kwargs.pop("hidden_states", None)
    weights = kwargs.pop("weights", None)
    if hidden_states is None and logits is None:
        raise VB200Error("hidden_states or logits must be provided.")
    sp_enabled = sp_group is not None and torch.distributed.get_world_size(sp_group) > 1
    if not sp_enabled:
        if shift_labels is None:
            labels = F.pad(labels, (0, 1), value=ignore_index)
            shift_labels = labels[..., 1:].contiguous()
    else:
        shift_labels = labels
    shift_labels = shift_labels.reshape(-1)
    if hidden_states is not None:
        hidden_states = hidden_states.reshape(-1, hidden_states.size(-1))
    if logits is not None:
        logits = logits.reshape(-1, vocab_size)
    loss, logits = cross_entropy_fn(logits, shift_labels, vocab_size, num_items_in_batch, ignore_index,
                                    hidden_states=hidden_states, weights=weights, **kwargs)
    if sp_enabled:
        loss = _ReduceLoss.apply(loss, (shift_labels != ignore_index).sum(), sp_group)
    return loss, logits, None
