"""Live per-launch timing of the package's kernels for bench.py (CUDA events on the launching stream).

``prof.ACTIVE = {"tag": [], ...}`` switches recording on for the listed tags; every ``with prof.span(tag, work):`` around a
kernel launch then appends ``(start_event, end_event, work)`` where ``work`` is the launch's algorithmic bytes or FLOPs
(SURVEY.md §8(d)). ``summarize`` turns the lists into totals after a synchronize. Off (None) by default: no events, no cost.
"""

from __future__ import annotations

import torch

ACTIVE: dict | None = None


class span:
    __slots__ = ("lst", "ev", "work")

    def __init__(self, tag: str, work: float = 0.0):
        self.lst = ACTIVE.get(tag) if ACTIVE is not None else None
        self.work = work

    def __enter__(self):
        if self.lst is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if self.lst is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self.lst.append((self.ev, end, self.work))
        return False


def summarize(rec: dict) -> dict:
    """{tag: {"launches", "ms", "work", "rate" (work per second)}} — call after torch.cuda.synchronize()."""
    out = {}
    for tag, lst in rec.items():
        if not lst:
            continue
        ms = sum(a.elapsed_time(b) for a, b, _ in lst)
        work = float(sum(w for _, _, w in lst))
        out[tag] = {"launches": len(lst), "ms": ms, "work": work, "rate": work / (ms * 1e-3) if ms > 0 else 0.0}
    return out
