"""FSDP2 parameter all-gather / gradient reduce-scatter on the NVLink pull kernels.

Drop-in boundary (SURVEY.md §8(b) "FSDP comm"): PyTorch's official hook
``FSDPModule.set_custom_all_gather(AllGather)`` / ``set_custom_reduce_scatter(ReduceScatter)``
(torch/distributed/fsdp/_fully_shard/_fully_shard.py:458-482; ABCs in _fsdp_api.py:56-127), which
VeOmni reaches through ``build_parallelize_model`` -> ``fully_shard``
(veomni/distributed/torch_parallelize.py:289-344).

* ``allocate`` hands FSDP2 a tensor that aliases this rank's symmetric region, so FSDP2's own copy-in
  (``fsdp::all_gather_copy_in`` / ``chunk_cat``) writes the local shard straight into peer-visible
  memory;
* ``__call__`` runs one kernel on FSDP2's all-gather / reduce-scatter stream (the current stream at
  call time): peers' shards are pulled over NVLink with 16-byte loads; the reduce-scatter sums in
  fp32 in rank order 0..N-1 (deterministic) and applies the AVG / pre-multiplied-SUM factor
  (``_get_gradient_divide_factors``, _fsdp_collectives.py:701-759) in the same pass.
"""

from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist
from torch.distributed.fsdp._fully_shard._fsdp_api import AllGather, ReduceScatter

from ._lib import VB200Error
from .symm import SymmetricMemory, get_symmetric_memory

CH_ALL_GATHER = 0
CH_REDUCE_SCATTER = 1


def _reduce_scale(op, world: int) -> float:
    """Scale applied to the fp32 sum for the ReduceOp FSDP2 passes."""
    ReduceOp = dist.ReduceOp
    if op == ReduceOp.SUM:
        return 1.0
    if op == ReduceOp.AVG:
        return 1.0 / world
    # PREMUL_SUM carries its factor in the op's pickled state: (RedOpType.PREMUL_SUM, factor)
    if op == ReduceOp.PREMUL_SUM:
        factor = op.__getstate__()[1]
        if isinstance(factor, torch.Tensor):
            factor = factor.item()
        return float(factor)
    raise VB200Error(f"unsupported reduce op for the B200 reduce-scatter: {op!r}")


class B200AllGather(AllGather):
    def __init__(self, symm: SymmetricMemory, num_ctas: int = 32):
        self.symm = symm
        self.num_ctas = num_ctas

    def allocate(self, size: Sequence[int], *, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        return self.symm.empty(tuple(int(s) for s in size), dtype, arena="fsdp_ag")

    def __call__(self, output_tensor, input_tensor, group, async_op: bool = False):
        if dist.get_world_size(group) != self.symm.world:
            raise VB200Error("all-gather group does not match the symmetric-memory group")
        if input_tensor.data_ptr() != output_tensor.data_ptr() + self.symm.rank * input_tensor.numel() * input_tensor.element_size():
            # FSDP2's copy-in always produces input = output[rank*n:(rank+1)*n]; keep the general case correct
            output_tensor.narrow(0, self.symm.rank * input_tensor.numel(), input_tensor.numel()).copy_(input_tensor)
        self.symm.all_gather_inplace(output_tensor, input_tensor.numel(), CH_ALL_GATHER, self.num_ctas)
        return None


class B200ReduceScatter(ReduceScatter):
    def __init__(self, symm: SymmetricMemory, num_ctas: int = 32):
        self.symm = symm
        self.num_ctas = num_ctas
        self._next_is_input = True

    def allocate(self, size: Sequence[int], *, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        # foreach_reduce (_fsdp_collectives.py:522-541) asks for the (N*chunk) input first and for the
        # (chunk) output second. Only the input must be peer-visible; the output backs the sharded
        # .grad for the rest of the step, so it comes from the ordinary caching allocator.
        shape = tuple(int(s) for s in size)
        if self._next_is_input:
            self._next_is_input = False
            return self.symm.empty(shape, dtype, arena="fsdp_rs")
        self._next_is_input = True
        return torch.empty(shape, dtype=dtype, device=device)

    def __call__(self, output_tensor, input_tensor, group, op, async_op: bool = False):
        world = dist.get_world_size(group)
        if world != self.symm.world:
            raise VB200Error("reduce-scatter group does not match the symmetric-memory group")
        if input_tensor.dtype != torch.float32:
            raise VB200Error(
                f"B200 reduce-scatter reduces in fp32 (VeOmni's default reduce_dtype, arguments_types.py:248-255); got {input_tensor.dtype}"
            )
        if not self.symm.contains(input_tensor):  # allocate() order assumption broken: stay correct
            staged = self.symm.empty(tuple(input_tensor.shape), input_tensor.dtype, arena="fsdp_rs")
            staged.copy_(input_tensor)
            input_tensor = staged
        self.symm.reduce_scatter_f32(input_tensor, output_tensor, _reduce_scale(op, world), CH_REDUCE_SCATTER,
                                     self.num_ctas)
        return None


def plan_fsdp_region(model: torch.nn.Module, world: int, param_bytes: int = 2, reduce_bytes: int = 4):
    """Size the symmetric region from the FSDP2 unit sizes of ``model``.

    Live all-gather outputs: the unit being copied out + the prefetched one; live reduce-scatter inputs:
    the one in flight + the previous one FSDP2 still holds. Returns (total_bytes, arena fractions).
    """
    from torch.distributed.fsdp import FSDPModule

    units = []
    for m in model.modules():
        if not isinstance(m, FSDPModule):
            continue
        nested = {id(p) for c in m.modules() if c is not m and isinstance(c, FSDPModule) for p in c.parameters()}
        numel = 0
        for p in m.parameters():
            if id(p) in nested:
                continue
            shape = tuple(p.shape)
            d0 = shape[0] if shape else 1
            rest = 1
            for s in shape[1:]:
                rest *= s
            numel += (d0 + world - 1) // world * world * rest
        units.append(numel)
    if not units:
        raise VB200Error("plan_fsdp_region: the model has no FSDP2 (fully_shard) modules")
    units.sort(reverse=True)
    top, second = units[0], (units[1] if len(units) > 1 else units[0])
    slack = 64 << 20
    ag = (top + 2 * second) * param_bytes + slack
    rs = (top + 2 * second) * reduce_bytes + slack
    misc = 256 << 20
    total = ag + rs + misc
    return total, {"fsdp_ag": ag / total, "fsdp_rs": rs / total, "misc": misc / total}


def install_fsdp_comm(model: torch.nn.Module, group: dist.ProcessGroup | None = None, symm: SymmetricMemory | None = None,
                      num_ctas: int = 32) -> SymmetricMemory:
    """Swap every FSDP2 unit of ``model`` onto the NVLink pull collectives.

    Call right after ``build_parallelize_model`` returns (veomni/trainer/base.py:387-404).
    """
    from torch.distributed.fsdp import FSDPModule

    if symm is None:
        g = group if group is not None else dist.group.WORLD
        total, arenas = plan_fsdp_region(model, dist.get_world_size(g))
        symm = get_symmetric_memory(g, total, arenas)
    ag, rs = B200AllGather(symm, num_ctas), B200ReduceScatter(symm, num_ctas)
    n = 0
    for m in model.modules():
        if isinstance(m, FSDPModule):
            m.set_custom_all_gather(ag)
            m.set_custom_reduce_scatter(rs)
            n += 1
    if n == 0:
        raise VB200Error("install_fsdp_comm: the model has no FSDP2 (fully_shard) modules")
    return symm
