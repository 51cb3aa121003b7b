"""Gradient-norm clipping for FSDP2 (+EP) — mirror of veomni/distributed/fsdp2/clip_grad_norm.py:21-153.

Scalars only: this stays on ``torch.distributed`` (NCCL) as ``north_star`` asks ("NCCL only for the outer DP
all-reduce"); no kernel of ours is involved.  Semantics:
* dense (non-expert) parameters: local p-norm^p over the DTensor shards, all-reduced over the FSDP group;
* expert parameters (tagged by ``ParallelPlan.apply`` with a ``Shard`` placement): all-reduced over the
  ``ep_fsdp`` group and then over the ``ep`` group (:124-137);
* one global clip coefficient for both groups (:145-151).
"""

from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch.distributed._tensor import DTensor

from .parallel_state import get_parallel_state


def _local(t: torch.Tensor) -> torch.Tensor:
    return t.to_local() if isinstance(t, DTensor) else t


def _reduce_group(params, norm_type: float, groups) -> torch.Tensor:
    dev = params[0].grad.device if params else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if math.isinf(norm_type):
        v = torch.zeros((), dtype=torch.float32, device=dev)
        for p in params:
            v = torch.maximum(v, _local(p.grad).detach().abs().max().float())
        for g in groups:
            if g is not None:
                dist.all_reduce(v, op=dist.ReduceOp.MAX, group=g)
        return v
    v = torch.zeros((), dtype=torch.float32, device=dev)
    if params:
        norms = torch._foreach_norm([_local(p.grad).detach() for p in params], norm_type)
        v = torch.stack([n.float() for n in norms]).pow(norm_type).sum()
    for g in groups:
        if g is not None:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=g)
    return v


@torch.no_grad()
def clip_grad_norm(model: torch.nn.Module, max_norm: float, norm_type: float = 2.0, error_if_nonfinite: bool = False,
                   foreach: bool | None = None) -> torch.Tensor:
    ps = get_parallel_state()
    expert, dense = [], []
    for p in model.parameters():
        if p.grad is None:
            continue
        info = getattr(p, "spec_info", None)
        (expert if info is not None and hasattr(info.placement, "dim") and ps.ep_enabled else dense).append(p)
    if not expert:
        total = torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], max_norm,
                                               norm_type=norm_type, error_if_nonfinite=error_if_nonfinite, foreach=foreach)
        return total.full_tensor() if isinstance(total, DTensor) else total
    fsdp_group = ps.fsdp_group if ps.device_mesh is not None and dist.is_initialized() else None
    ep_fsdp_group = ps.ep_fsdp_device_mesh["ep_fsdp"].get_group() if ps.ep_fsdp_device_mesh is not None else None
    d = _reduce_group(dense, norm_type, [fsdp_group])
    e = _reduce_group(expert, norm_type, [ep_fsdp_group, ps.ep_group])
    total = torch.maximum(d, e) if math.isinf(norm_type) else (d + e).pow(1.0 / norm_type)
    if error_if_nonfinite and not torch.isfinite(total):
        raise RuntimeError(f"The total norm of order {norm_type} for gradients is non-finite")
    torch.nn.utils.clip_grads_with_norm_(expert, max_norm, total, foreach=foreach)
    torch.nn.utils.clip_grads_with_norm_(dense, max_norm, total, foreach=foreach)
    return total
