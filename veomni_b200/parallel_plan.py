"""ParallelPlan — expert-parallel weight slicing; mirror of veomni/distributed/parallel_plan.py:45-213.

A plan maps fqn patterns (``*`` matches one dotted component, as ``check_fqn_match``,
veomni/distributed/utils.py:109-122) to ``Shard(dim)``.  ``apply`` replaces every matching parameter by this
EP rank's contiguous slice along ``dim`` (rank r owns experts ``[r*E/EP, (r+1)*E/EP)``, the layout
``DTensor.redistribute(Replicate -> Shard(0)).to_local()`` produces at :77-85) and tags parameters with
``spec_info`` so later stages (FSDP wrapping on dim 1, grad-norm clipping) can tell expert parameters apart.
"""

from __future__ import annotations

import re
from dataclasses import dataclass

import torch
from torch import nn
from torch.distributed._tensor import Replicate, Shard


@dataclass
class SpecInfo:
    para_name: str
    placement: object
    fqn: str
    ep_size: int = 1
    ep_rank: int = 0


def check_fqn_match(pattern: str, fqn: str) -> bool:
    rx = "^" + re.escape(pattern).replace(r"\*", r"[^.]+") + "$"
    return re.match(rx, fqn) is not None


def _set_by_path(model: nn.Module, fqn: str, value: nn.Parameter) -> None:
    parts = fqn.split(".")
    mod = model
    for p in parts[:-1]:
        mod = getattr(mod, p)
    setattr(mod, parts[-1], value)


class ParallelPlan:
    def __init__(self, extra_parallel_plan: dict[str, dict[str, Shard]]):
        self.extra_parallel_plan = extra_parallel_plan

    def shard_tensor(self, tensor: torch.Tensor, fqn: str, ep_size: int, ep_rank: int) -> torch.Tensor:
        """Slice a full (checkpoint) tensor for this rank — used when loading weights."""
        for plan in self.extra_parallel_plan.values():
            for pattern, shard in plan.items():
                if check_fqn_match(pattern, fqn):
                    n = tensor.size(shard.dim)
                    if n % ep_size:
                        raise AssertionError(f"{fqn}: dim {shard.dim} of size {n} not divisible by {ep_size}")
                    return tensor.narrow(shard.dim, ep_rank * (n // ep_size), n // ep_size).contiguous()
        return tensor

    def apply(self, model: nn.Module, ep_size: int, ep_rank: int, para_name: str = "ep") -> dict[str, SpecInfo]:
        plan = self.extra_parallel_plan.get(para_name, {})
        out: dict[str, SpecInfo] = {}
        for fqn, param in list(model.named_parameters()):
            matched = None
            for pattern, shard in plan.items():
                if check_fqn_match(pattern, fqn):
                    matched = shard
                    break
            if matched is not None and ep_size > 1:
                local = nn.Parameter(self.shard_tensor(param.data, fqn, ep_size, ep_rank), requires_grad=param.requires_grad)
                info = SpecInfo(para_name, matched, fqn, ep_size, ep_rank)
                local.spec_info = info
                _set_by_path(model, fqn, local)
            else:
                info = SpecInfo(para_name, Replicate(), fqn, ep_size, ep_rank)
                param.spec_info = info
            out[fqn] = info
        return out


def qwen3_moe_parallel_plan() -> ParallelPlan:
    """veomni/models/transformers/qwen3_moe/parallel_plan.py:6-16."""
    return ParallelPlan({"ep": {
        "model.layers.*.mlp.experts.gate_up_proj": Shard(0),
        "model.layers.*.mlp.experts.down_proj": Shard(0),
    }})
