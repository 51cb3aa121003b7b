"""tools/stream_gaps.py on a synthetic timeline (no GPU): stream statistics, gap detection and classification."""
import torch

from tools import stream_gaps


class _Ev:
    def __init__(self, s, e, r, n):
        self._s, self._e, self._r, self._n = s, e, r, n

    def device_type(self):
        return torch.autograd.DeviceType.CUDA

    def duration_ns(self):
        return self._e - self._s

    def start_ns(self):
        return self._s

    def end_ns(self):
        return self._e

    def device_resource_id(self):
        return self._r

    def name(self):
        return self._n


class _TP:
    def __init__(self, evs):
        outer = self

        class _K:
            def events(self_inner):
                return evs

        class _P:
            kineto_results = _K()

        self.profiler = _P()


def test_gaps_are_found_and_classified():
    us = 1000
    evs = [
        _Ev(0, 100 * us, 7, "gemm_a"),
        _Ev(400 * us, 500 * us, 7, "gemm_b"),      # 300 us gap, covered by an all-gather on stream 9
        _Ev(900 * us, 1000 * us, 7, "gemm_c"),     # 400 us gap with nothing running
        _Ev(90 * us, 390 * us, 9, "void vb::allgather_scatter_kernel(...)"),
    ]
    txt = stream_gaps.report(_TP(evs), min_gap_us=30)
    assert "compute stream 7: 2 gaps" in txt
    assert "0.70 ms idle" in txt
    assert "while allgather_scatter_kernel" in txt
    assert "nothing running on any stream" in txt


def test_empty_profile():
    assert "no device events" in stream_gaps.report(_TP([]))
