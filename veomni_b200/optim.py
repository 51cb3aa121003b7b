"""AdamW on the multi-tensor kernel (EXPERIMENTAL: written in round 1, not yet validated on hardware — nothing in the
default path uses it; ``bench.py --b200-adamw`` and the gated test in tests/test_ops_gpu.py exercise it).

The reference builds ``torch.optim.AdamW(..., fused=True)`` over the fp32 DTensor shards
(veomni/optim/optimizer.py:261-328; SURVEY.md §8(f)4 lists the optimizer step as "next"). PyTorch's fused kernel moves
the 28 bytes per parameter at ~4.5 TB/s on the ~400 shards of Qwen3-8B (51 ms of the 339 ms single-GPU step); this
class runs the same arithmetic as one launch over a device table of (param, grad, exp_avg, exp_avg_sq) entries, and can
fold a gradient-clip coefficient in (``grad_scale``) instead of a separate scaling pass.

Same hyper-parameters, same state keys (``step``, ``exp_avg``, ``exp_avg_sq``) as ``torch.optim.AdamW``; no amsgrad,
maximize, capturable or differentiable modes; fp32 CUDA parameters only (anything else raises).
"""

from __future__ import annotations

import math

import torch
from torch.distributed._tensor import DTensor

from . import _lib
from ._lib import VB200Error, check, stream_ptr

_ENTRY = 1 << 20


def _local(t: torch.Tensor) -> torch.Tensor:
    return t.to_local() if isinstance(t, DTensor) else t


class B200AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables: dict = {}

    def _table(self, gi: int, tensors: list[list[torch.Tensor]]):
        """Device pointer table for group ``gi``: rows = (param, grad, exp_avg, exp_avg_sq) pointers, last row numels;
        rebuilt only when a pointer changed (gradients come back at the same addresses from the caching allocator)."""
        key = tuple(t.data_ptr() for ts in tensors for t in ts)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        rows = [[], [], [], [], []]
        for p, g, m, v in zip(*tensors):
            n = p.numel()
            for o in range(0, n, _ENTRY):
                for r, t in zip(rows, (p, g, m, v)):
                    r.append(t.data_ptr() + o * 4)
                rows[4].append(min(_ENTRY, n - o))
        dev = tensors[0][0].device
        tab = torch.tensor(rows, dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
        self._tables[gi] = (key, tab, len(rows[0]))
        return tab, len(rows[0])

    @torch.no_grad()
    def step(self, closure=None, grad_scale: torch.Tensor | None = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            ps, gs, ms, vs = [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                pl, gl = _local(p), _local(p.grad)
                if not (pl.is_cuda and pl.dtype == torch.float32 and gl.dtype == torch.float32 and pl.is_contiguous()
                        and gl.is_contiguous()):
                    raise VB200Error("B200AdamW: contiguous fp32 CUDA parameters and gradients only")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(pl)
                    st["exp_avg_sq"] = torch.zeros_like(pl)
                st["step"] += 1
                ps.append(pl); gs.append(gl); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
            if not ps:
                continue
            steps = {self.state[p]["step"] for p in group["params"] if p.grad is not None}
            if len(steps) != 1:
                raise VB200Error("B200AdamW: parameters of one group must share their step count")
            t = steps.pop()
            b1, b2 = group["betas"]
            tab, n = self._table(gi, [ps, gs, ms, vs])
            gsd = None
            if grad_scale is not None:
                gsd = grad_scale.to(device=ps[0].device, dtype=torch.float32).reshape(1).contiguous()
            with torch.cuda.device(ps[0].device):
                check(lib.vb200_multi_adamw(tab[0].data_ptr(), tab[1].data_ptr(), tab[2].data_ptr(), tab[3].data_ptr(),
                                            tab[4].data_ptr(), n, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                            float(group["weight_decay"]), 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t),
                                            gsd.data_ptr() if gsd is not None else None, stream_ptr()), "vb200_multi_adamw")
        return loss
