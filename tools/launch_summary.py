"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with the live numbers, not absolutes.
"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("us", "usecond"):
            v *= 1e3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e6
        name = re.sub(r"<.*", "", r[ki]).split("(")[0].strip()
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for k, v in tot.items() if k.startswith("vb::") or "vb::" in k)
    print(f"# {path}: {sum(cnt.values())} launches, {total/1e6:.2f} ms kernel time, veomni_b200 kernels {100*ours/total:.1f}% of it")
    print(f"{'share':>7} {'ms':>9} {'count':>6}  kernel")
    for k, v in tot.most_common(40):
        print(f"{100*v/total:6.2f}% {v/1e6:9.3f} {cnt[k]:6d}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
