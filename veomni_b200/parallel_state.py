"""Parallel topology for the hot path — mirror of veomni/distributed/parallel_state.py.

Same mesh construction as ``init_parallel_state`` (:485-623): dims ``[pp, dp_replicate, dp_shard, ulysses, cp, tp]``
keeping only sizes > 1 (always ``dp_shard``), flattened views ``dp``, ``dp_shard_sp`` (the FSDP mesh — Ulysses
ranks are also FSDP shard ranks, :87,95-96), ``dp_sp`` and ``sp``; a separate 2-D ``(ep, ep_fsdp)`` mesh over the
same ranks for expert parallelism (``init_para_mesh_matrix`` :50-73: EP ranks are consecutive unless placed
innermost).  TP / PP / CP are rejected exactly like the reference (arguments_types.py:561-562,
parallel_state.py:98-99).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.distributed as dist
from torch.distributed.device_mesh import DeviceMesh, init_device_mesh

_STATE: "ParallelState | None" = None


def init_para_mesh_matrix(para_size: int, para_fsdp_size: int, para_outside: bool = False) -> torch.Tensor:
    """Rank matrix [para, para_fsdp] (reference :50-73)."""
    n = math.prod((para_size, para_fsdp_size))
    if para_outside:
        return torch.arange(n, dtype=torch.int).view(para_size, para_fsdp_size)
    return torch.arange(n, dtype=torch.int).view(para_fsdp_size, para_size).transpose(0, 1)


@dataclass(frozen=True)
class ParallelState:
    dp_size: int = 1
    dp_replicate_size: int = 1
    dp_shard_size: int = 1
    ulysses_size: int = 1
    ep_size: int = 1
    dp_mode: str = "fsdp2"
    device_type: str = "cuda"
    device_mesh: DeviceMesh | None = None
    ep_fsdp_device_mesh: DeviceMesh | None = None
    async_enabled: bool = False
    _cache: dict = field(default_factory=dict, compare=False, repr=False)

    # -- data / fsdp --------------------------------------------------------------------------
    @property
    def world_size(self) -> int:
        return dist.get_world_size() if dist.is_initialized() else 1

    @property
    def fsdp_enabled(self) -> bool:
        return self.dp_mode == "fsdp2" and self.world_size > 1 or self.dp_shard_size >= 1 and self.device_mesh is not None

    @property
    def fsdp_mesh(self) -> DeviceMesh:
        """Reference :253-268: a 2-D (replicate, shard) mesh under HSDP, the flattened shard mesh otherwise."""
        if self.dp_replicate_size > 1:
            if self.ulysses_size > 1 and self.dp_shard_size > 1:
                return self.device_mesh["dp_replicate", "dp_shard_sp"]
            return self.device_mesh["dp_replicate", "dp_shard"]
        return self.device_mesh["dp_shard_sp"]

    @property
    def fsdp_group(self):
        """The group parameters are sharded over (all-gather / reduce-scatter group)."""
        m = self.fsdp_mesh
        return m.get_group(m.ndim - 1) if m.ndim > 1 else m.get_group()

    @property
    def dp_group(self):
        return self.device_mesh["dp"].get_group()

    # -- sequence parallel --------------------------------------------------------------------
    @property
    def ulysses_enabled(self) -> bool:
        return self.ulysses_size > 1

    @property
    def sp_enabled(self) -> bool:
        return self.ulysses_size > 1

    @property
    def ulysses_group(self):
        return self.device_mesh["ulysses"].get_group() if self.ulysses_enabled else None

    @property
    def ulysses_rank(self) -> int:
        return dist.get_rank(self.ulysses_group) if self.ulysses_enabled else 0

    # -- expert parallel ----------------------------------------------------------------------
    @property
    def ep_enabled(self) -> bool:
        return self.ep_size > 1

    @property
    def ep_group(self):
        return self.ep_fsdp_device_mesh["ep"].get_group() if self.ep_enabled else None

    @property
    def ep_rank(self) -> int:
        return dist.get_rank(self.ep_group) if self.ep_enabled else 0

    @property
    def ep_fsdp_size(self) -> int:
        return self.world_size // self.ep_size

    def extra_parallel_gradient_divide_factor(self, para: str = "ep") -> int:
        """Expert grads are averaged over the whole world (reference :381-389, torch_parallelize.py:306-313)."""
        return self.world_size


def init_parallel_state(dp_size: int = 1, dp_replicate_size: int = 1, dp_shard_size: int = 1, tp_size: int = 1,
                        pp_size: int = 1, cp_size: int = 1, ulysses_size: int = 1, dp_mode: str = "fsdp2",
                        device_type: str | None = None, ep_size: int = 1, ep_outside: bool = False,
                        async_enabled: bool = False) -> ParallelState:
    global _STATE
    if _STATE is not None:
        return _STATE
    if tp_size != 1 or pp_size != 1:
        raise AssertionError("tp_size and pp_size must be 1 (not implemented in the reference either)")
    if cp_size > 1:
        raise NotImplementedError("Ring attention is not supported yet.")
    world = dist.get_world_size()
    if dp_size * ulysses_size != world:
        raise ValueError("The product of parallel sizes should be equal to the world size.")
    if dp_size > 1 and dp_shard_size == 1 and dp_replicate_size == 1:
        dp_shard_size = dp_size
    if dp_replicate_size * dp_shard_size != dp_size:
        raise ValueError(f"The product of dp_replicate_size: {dp_replicate_size} and dp_shard_size: {dp_shard_size} "
                         f"should be equal to dp_size: {dp_size}.")
    if device_type is None:
        device_type = "cuda" if torch.cuda.is_available() else "cpu"
    shape, names = [], []
    for d, n in zip([dp_replicate_size, dp_shard_size, ulysses_size], ["dp_replicate", "dp_shard", "ulysses"]):
        if d > 1 or n == "dp_shard":
            shape.append(d)
            names.append(n)
    mesh = init_device_mesh(device_type, tuple(shape), mesh_dim_names=tuple(names))
    dp_names = [n for n in names if n in ("dp_replicate", "dp_shard")]
    shard_sp = [n for n in names if n in ("dp_shard", "ulysses")]
    mesh[tuple(dp_names)]._flatten(mesh_dim_name="dp")
    mesh[tuple(shard_sp)]._flatten(mesh_dim_name="dp_shard_sp")
    mesh[tuple(names)]._flatten(mesh_dim_name="dp_sp")
    if ulysses_size > 1:
        mesh[("ulysses",)]._flatten(mesh_dim_name="sp")
    ep_mesh = None
    if ep_size > 1:
        if world % ep_size:
            raise AssertionError("ep_size must be a factor of world_size")
        ep_mesh = DeviceMesh(device_type, init_para_mesh_matrix(ep_size, world // ep_size, ep_outside),
                             mesh_dim_names=("ep", "ep_fsdp"))
    _STATE = ParallelState(dp_size, dp_replicate_size, dp_shard_size, ulysses_size, ep_size, dp_mode, device_type, mesh,
                           ep_mesh, async_enabled)
    return _STATE


def get_parallel_state() -> ParallelState:
    return _STATE if _STATE is not None else ParallelState()


def reset_parallel_state() -> None:
    global _STATE
    _STATE = None
