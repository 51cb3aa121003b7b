"""Step-level roofline of the in-scope kernels from an ncu launch list of `bench.py` (BASELINE.md §3 reporting rule):
sum of algorithmic minimum times of our kernels / sum of their measured times, per kernel and in total.

    python tools/step_roofline.py profiles/r01_launches_bench_n1.csv.gz > profiles/r01_step_roofline.txt

Algorithmic work per launch at the Qwen3-8B shapes of the bench (T = 4096 tokens, H = 4096, I = 12288, 32/8 heads,
D = 128, V = 151936); peaks from MEASURED_PEAKS.json when present (else the profiling guide's
fallbacks). Times under ncu are cold-cache and serialised, so the fractions are pessimistic for the HBM kernels.
"""
import collections
import csv
import gzip
import json
import re
import sys
from pathlib import Path

T, H, I, HQ, HK, D, V = 4096, 4096, 12288, 32, 8, 128, 151936
FWD_ATTN = 4 * T * T * D * HQ / 2  # causal forward FLOPs per launch
P8 = 8.191e9  # parameters of Qwen3-8B (untied)
# (regex on the kernel name incl. template arguments, "hbm" | "tensor", algorithmic bytes or FLOPs per launch)
WORK = [
    (r"vb::add_rmsnorm_fwd_kernel<\d+, (false|0)>", "hbm", 2 * T * H * 2),          # plain RMSNorm forward (register-resident)
    (r"vb::add_rmsnorm_fwd_kernel<\d+, (true|1)>", "hbm", 4 * T * H * 2),           # fused residual add + RMSNorm
    (r"vb::rmsnorm_fwd_bulk_kernel", "hbm", 2 * T * H * 2),
    (r"vb::rmsnorm_bwd_ring_kernel<\d+, (false|0)>", "hbm", 3 * T * H * 2),
    (r"vb::rmsnorm_bwd_ring_kernel<\d+, (true|1)>", "hbm", 4 * T * H * 2),
    (r"vb::rmsnorm_bwd_wide_kernel", "hbm", 3 * T * H * 2),
    (r"vb::rmsnorm_bwd_wide_add_kernel", "hbm", 4 * T * H * 2),
    (r"vb::colsum2?_kernel", "hbm", 0),                                               # second pass of the dw reduction: overhead
    (r"vb::qknorm_rope_fwd_kernel", "hbm", 2 * T * (HQ + HK) * D * 2),
    (r"vb::qknorm_rope_bwd_kernel", "hbm", 3 * T * (HQ + HK) * D * 2),
    (r"vb::swiglu_fwd_kernel", "hbm", 3 * T * I * 2),
    (r"vb::swiglu_bwd_kernel", "hbm", 5 * T * I * 2),
    (r"vb::attn_fwd_tc_kernel", "tensor", FWD_ATTN),
    (r"vb::attn_bwd_dq_(tc|n128)_kernel", "tensor", FWD_ATTN * 2.5 * 0.4),
    (r"vb::attn_bwd_dkdv_tc_kernel", "tensor", FWD_ATTN * 2.5 * 0.6),
    (r"vb::attn_bwd_delta_kernel", "hbm", 2 * T * HQ * D * 2),
    (r"vb::cross_entropy_kernel", "hbm", 2 * 1024 * V * 2),
    (r"vb::multi_adamw_kernel<__nv_bfloat16, (true|1)>", "hbm", 28 * P8),           # fp32 master + 2 moments read+write (24), bf16 grad read (2), bf16 copy write (2)
    (r"vb::multi_sumsq_kernel", "hbm", 2 * P8),                                       # bf16 gradients read once
]


def peaks():
    p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d["hbm_gbs"] * 1e9, d.get("bf16_tflops_sustained", d["bf16_tflops"]) * 1e12, "measured"
    return 6650e9, 1590e12, "fallback"


def main(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(r[ui], 1.0)
        name = re.sub(r"^void ", "", r[ki].split("(")[0].strip())
        tot[name] += v * 1e-9
        cnt[name] += 1
    hbm, tensor, how = peaks()
    print(f"# {path}: step-level roofline of the in-scope kernels ({how} peaks: {hbm/1e9:.0f} GB/s, {tensor/1e12:.0f} TF/s)")
    print(f"{'kernel':<44} {'launches':>8} {'measured ms':>12} {'algorithmic ms':>15} {'fraction':>9}")
    s_min = s_meas = 0.0
    seen = set()
    o_min = o_meas = 0.0  # optimizer-side kernels (AdamW, gradient norm), reported apart as well
    for pat, bound, work in WORK:
        names = [n for n in cnt if re.match(pat, n)]
        n_l = sum(cnt[n] for n in names)
        if not n_l:
            continue
        seen.update(names)
        t_meas = sum(tot[n] for n in names)
        tmin = n_l * work / (hbm if bound == "hbm" else tensor)
        s_min += tmin
        s_meas += t_meas
        if "multi_" in pat:
            o_min += tmin
            o_meas += t_meas
        label = re.sub(r"\\d\+|\(|\)|\\", "", pat)[4:]
        print(f"{label[:44]:<44} {n_l:>8d} {t_meas*1e3:>12.3f} {tmin*1e3:>15.3f} {tmin/t_meas:>9.3f}")
    other = {n: tot[n] for n in cnt if n.startswith("vb::") and n not in seen}
    for n, t in sorted(other.items(), key=lambda kv: -kv[1])[:8]:
        print(f"# not in the table: {n[:70]} {cnt[n]} launches, {t*1e3:.3f} ms")
    print(f"{'TOTAL (in-scope kernels)':<44} {'':>8} {s_meas*1e3:>12.3f} {s_min*1e3:>15.3f} {s_min/s_meas:>9.3f}")
    if o_meas and s_meas > o_meas:
        print(f"{'TOTAL without the optimizer kernels':<44} {'':>8} {(s_meas-o_meas)*1e3:>12.3f} {(s_min-o_min)*1e3:>15.3f} "
              f"{(s_min-o_min)/(s_meas-o_meas):>9.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
