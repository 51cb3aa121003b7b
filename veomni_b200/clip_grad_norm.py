"""Gradient-norm clipping for FSDP2 (+EP) — mirror of veomni/distributed/fsdp2/clip_grad_norm.py:21-153.

Semantics (unchanged from the reference):
* dense (non-expert) parameters: local p-norm^p over the DTensor shards, all-reduced over the FSDP group;
* expert parameters (tagged by ``ParallelPlan.apply`` with a ``Shard`` placement): all-reduced over the
  ``ep_fsdp`` group and then over the ``ep`` group (:124-137);
* one global clip coefficient ``max_norm / (total + 1e-6)`` clamped to 1 for both groups (:145-151).

The cross-rank part is a scalar all-reduce and stays on ``torch.distributed`` (NCCL), as ``north_star`` asks. The two
memory passes — sum of squares of every local shard, and the in-place scale — are the multi-tensor kernels of
``csrc/multi_tensor.cu`` for the L2 norm of CUDA fp32 / bf16 gradients (PyTorch's ``_foreach_norm`` /
``_foreach_mul_`` reach ~1 / 2.4 TB/s on the ~400 shards of Qwen3-8B). The coefficient never visits the host.
Other norm types keep PyTorch's foreach ops (library code, GPU).
"""

from __future__ import annotations

import math

import torch
import torch.distributed as dist
from torch.distributed._tensor import DTensor

from .parallel_state import get_parallel_state


def _local(t: torch.Tensor) -> torch.Tensor:
    return t.to_local() if isinstance(t, DTensor) else t


_ENTRY = 1 << 20  # elements per table entry (csrc/multi_tensor.cu)
_DT = {torch.bfloat16: 0, torch.float32: 1}
_tables: dict = {}


def _table(tensors: list[torch.Tensor]):
    """Device (pointer, numel) table of the tensors split into entries of at most 2^20 elements; cached on the exact
    pointer list (the caching allocator hands the gradient shards the same blocks every step)."""
    key = tuple((t.data_ptr(), t.numel()) for t in tensors)
    dev = tensors[0].device
    hit = _tables.get((dev, len(key)))
    if hit is not None and hit[0] == key:
        return hit[1], hit[2], hit[3]
    esz = tensors[0].element_size()
    ptrs, nums = [], []
    for ptr, n in key:
        for o in range(0, n, _ENTRY):
            ptrs.append(ptr + o * esz)
            nums.append(min(_ENTRY, n - o))
    host = torch.tensor([ptrs, nums], dtype=torch.int64).pin_memory()
    devt = host.to(dev, non_blocking=True)
    _tables[(dev, len(key))] = (key, devt[0], devt[1], len(ptrs))
    return devt[0], devt[1], len(ptrs)


def multi_sumsq(tensors: list[torch.Tensor]) -> torch.Tensor:
    """fp32 device scalar sum(x^2) over all tensors (contiguous CUDA tensors of one dtype, fp32 or bf16)."""
    from . import _lib
    from ._lib import VB200Error, check, stream_ptr

    if not tensors:
        raise VB200Error("multi_sumsq: empty tensor list")
    t0 = tensors[0]
    if not all(t.is_cuda and t.dtype == t0.dtype and t.is_contiguous() for t in tensors) or t0.dtype not in _DT:
        raise VB200Error("multi_sumsq: contiguous CUDA tensors of one dtype (fp32 or bf16) expected")
    lib = _lib.load()
    ptrs, nums, n = _table(tensors)
    out = torch.empty(1, dtype=torch.float32, device=t0.device)
    partials = torch.empty(max(1, lib.vb200_multi_sumsq_partials(n)), dtype=torch.float32, device=t0.device)
    with torch.cuda.device(t0.device):
        check(lib.vb200_multi_sumsq(ptrs.data_ptr(), nums.data_ptr(), n, _DT[t0.dtype], partials.data_ptr(), out.data_ptr(),
                                    None, stream_ptr()), "vb200_multi_sumsq")
    return out[0]


def multi_scale_(tensors: list[torch.Tensor], coef: torch.Tensor) -> None:
    """x *= coef (a CUDA fp32 scalar tensor) for every tensor, in place; a coefficient of exactly 1 is a no-op."""
    from . import _lib
    from ._lib import check, stream_ptr

    if not tensors:
        return
    t0 = tensors[0]
    lib = _lib.load()
    ptrs, nums, n = _table(tensors)
    coef = coef.to(device=t0.device, dtype=torch.float32).reshape(1).contiguous()
    with torch.cuda.device(t0.device):
        check(lib.vb200_multi_scale(ptrs.data_ptr(), nums.data_ptr(), n, _DT[t0.dtype], coef.data_ptr(), stream_ptr()),
              "vb200_multi_scale")


def _kernel_path_ok(grads: list[torch.Tensor], norm_type: float) -> bool:
    return (norm_type == 2.0 and len(grads) > 0 and all(g.is_cuda and g.is_contiguous() for g in grads)
            and len({g.dtype for g in grads}) == 1 and grads[0].dtype in _DT)


def _shard_groups(t: torch.Tensor) -> tuple:
    """Process groups a DTensor gradient is sharded over (its partial sums must be added across them)."""
    if not isinstance(t, DTensor):
        return ()
    return tuple(t.device_mesh.get_group(i) for i, pl in enumerate(t.placements) if pl.is_shard())


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
        grads = [_local(p.grad).detach() for p in params]
        if _kernel_path_ok(grads, norm_type):
            v = multi_sumsq(grads)
        else:
            norms = torch._foreach_norm(grads, norm_type)
            v = torch.stack([n.float() for n in norms]).pow(norm_type).sum()
    for g in groups:
        if g is not None:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=g)
    return v


@torch.no_grad()
def clip_grad_norm(model: torch.nn.Module, max_norm: float, norm_type: float = 2.0, error_if_nonfinite: bool = False,
                   foreach: bool | None = None, return_coef: bool = False):
    """Same contract as the reference. ``return_coef=True`` (ours): the gradients are left untouched and
    ``(total_norm, coef)`` is returned, ``coef`` a device scalar to hand to ``B200AdamW.step(grad_scale=coef)`` — the
    scaling pass over every gradient folds into the optimizer kernel. Only the plain FSDP2 L2 path supports it."""
    ps = get_parallel_state()
    expert, dense = [], []
    for p in model.parameters():
        if p.grad is None:
            continue
        info = getattr(p, "spec_info", None)
        (expert if info is not None and hasattr(info.placement, "dim") and ps.ep_enabled else dense).append(p)
    if not expert:
        grads = [_local(p.grad) for p in dense]
        if not _kernel_path_ok(grads, norm_type):
            if return_coef:
                raise NotImplementedError("clip_grad_norm(return_coef=True) needs contiguous CUDA fp32 / bf16 gradients and the L2 norm")
            total = torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], max_norm,
                                                   norm_type=norm_type, error_if_nonfinite=error_if_nonfinite, foreach=foreach)
            return total.full_tensor() if isinstance(total, DTensor) else total
        # bucket by the groups each gradient is sharded over (one bucket for plain FSDP2)
        buckets: dict = {}
        for p, g in zip(dense, grads):
            buckets.setdefault(_shard_groups(p.grad), []).append(g)
        total_sq = None
        for groups, gs in buckets.items():
            v = multi_sumsq(gs)
            for grp in groups:
                if dist.get_world_size(grp) > 1:
                    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=grp)
            total_sq = v if total_sq is None else total_sq + v
        total = total_sq.sqrt()
        if error_if_nonfinite and not torch.isfinite(total):
            raise RuntimeError(f"The total norm of order {norm_type} for gradients is non-finite")
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        if return_coef:
            return total, coef
        multi_scale_(grads, coef)
        return total
    if return_coef:
        raise NotImplementedError("clip_grad_norm(return_coef=True) covers the non-EP L2 path only")
    fsdp_group = ps.fsdp_group if ps.device_mesh is not None and dist.is_initialized() else None
    ep_fsdp_group = ps.ep_fsdp_device_mesh["ep_fsdp"].get_group() if ps.ep_fsdp_device_mesh is not None else None
    d = _reduce_group(dense, norm_type, [fsdp_group])
    e = _reduce_group(expert, norm_type, [ep_fsdp_group, ps.ep_group])
    total = torch.maximum(d, e) if math.isinf(norm_type) else (d + e).pow(1.0 / norm_type)
    if error_if_nonfinite and not torch.isfinite(total):
        raise RuntimeError(f"The total norm of order {norm_type} for gradients is non-finite")
    for group in (expert, dense):
        grads = [_local(p.grad) for p in group]
        if _kernel_path_ok(grads, 2.0):
            multi_scale_(grads, torch.clamp(max_norm / (total + 1e-6), max=1.0))
        else:
            torch.nn.utils.clip_grads_with_norm_(group, max_norm, total, foreach=foreach)
    return total
