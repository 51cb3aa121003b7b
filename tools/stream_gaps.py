"""Where does a profiled step idle?  Per-stream busy time and the largest gaps of the compute stream.

``report(tp)`` takes a finished ``torch.profiler.profile`` (CUDA activity) of ONE step and returns a text report:
* every CUDA stream with its kernel count, busy time and span;
* the compute stream (the one with the most kernel time that is not a collective's stream) — its idle gaps sorted by
  length, each with the kernel before it, the kernel after it and the collectives that were running on other streams
  while it idled (a gap that ends when an all-gather ends is an exposed all-gather; a gap with nothing running anywhere
  is the host not keeping up).

Used by ``bench.py --torch-profile`` (appended to the kernel table) so that the residual of the multi-GPU step is named
from a measurement instead of guessed.
"""
from __future__ import annotations

from collections import defaultdict

COMM = ("allgather", "reduce_scatter", "all_to_all", "chunk_pull", "nccl", "barrier_kernel", "fsdp_pack")


def _short(name: str, n: int = 70) -> str:
    name = name.replace("void ", "").replace("vb::", "")
    return name if len(name) <= n else name[: n - 1] + "…"


def report(tp, top: int = 24, min_gap_us: float = 30.0) -> str:
    import torch

    evs = []
    for e in tp.profiler.kineto_results.events():
        if e.device_type() != torch.autograd.DeviceType.CUDA or e.duration_ns() <= 0:
            continue
        evs.append((e.start_ns() / 1e3, e.end_ns() / 1e3, int(e.device_resource_id()), e.name()))
    if not evs:
        return "stream_gaps: no device events\n"
    t0 = min(s for s, _e, _r, _n in evs)
    t1 = max(e for _s, e, _r, _n in evs)
    by_stream = defaultdict(list)
    for s, e, r, n in evs:
        by_stream[r].append((s, e, n))
    lines = [f"# step span {(t1 - t0) / 1e3:.2f} ms, {len(evs)} device events on {len(by_stream)} streams"]

    def busy(iv):
        iv = sorted(iv)
        tot, cs, ce = 0.0, None, None
        for s, e, _n in iv:
            if cs is None:
                cs, ce = s, e
            elif s <= ce:
                ce = max(ce, e)
            else:
                tot += ce - cs
                cs, ce = s, e
        return tot + (ce - cs if cs is not None else 0.0)

    stats = {}
    for r, iv in by_stream.items():
        comm_t = sum(e - s for s, e, n in iv if any(c in n.lower() for c in COMM))
        stats[r] = (busy(iv), comm_t, len(iv))
    for r, (b, c, k) in sorted(stats.items(), key=lambda t: -t[1][0]):
        lines.append(f"stream {r:>4}: {k:5d} kernels, busy {b / 1e3:8.2f} ms ({100 * b / (t1 - t0):5.1f} % of the span), of which collectives {c / 1e3:7.2f} ms")
    compute = max(stats, key=lambda r: stats[r][0] - stats[r][1])
    iv = sorted(by_stream[compute])
    comm_iv = sorted((s, e, n) for r, v in by_stream.items() if r != compute for s, e, n in v if any(c in n.lower() for c in COMM))
    gaps = []
    end, prev = iv[0][1], iv[0][2]
    for s, e, n in iv[1:]:
        if s - end >= min_gap_us:
            gaps.append((s - end, end, s, prev, n))
        if e > end:
            end, prev = e, n
    idle = sum(g[0] for g in gaps)
    lines.append(f"compute stream {compute}: {len(gaps)} gaps >= {min_gap_us:.0f} us, {idle / 1e3:.2f} ms idle in total")
    # classify every gap by what else was running
    classes = defaultdict(float)
    detail = []
    for g, gs, ge, before, after in gaps:
        running = [(n, max(gs, s), min(ge, e)) for s, e, n in comm_iv if s < ge and e > gs]
        cover = defaultdict(float)
        for n, a, b in running:
            cover[_short(n, 40)] += b - a
        if cover:
            top_n, top_t = max(cover.items(), key=lambda t: t[1])
            key = f"while {top_n} ran" if top_t > 0.5 * g else f"partly under {top_n}"
        else:
            key = "nothing running on any stream (host / dependency on the host)"
        classes[key] += g
        detail.append((g, gs - t0, before, after, key))
    lines.append("idle time by what was running on the other streams:")
    for k, v in sorted(classes.items(), key=lambda t: -t[1]):
        lines.append(f"  {v / 1e3:8.2f} ms  {k}")
    lines.append(f"largest {top} gaps (us | at ms | after kernel -> before kernel | other streams):")
    for g, at, before, after, key in sorted(detail, reverse=True)[:top]:
        lines.append(f"  {g:8.0f} | {at / 1e3:8.2f} | {_short(before, 48)} -> {_short(after, 48)} | {key}")
    return "\n".join(lines) + "\n"
