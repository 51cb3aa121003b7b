"""Per-instruction stall reasons from an .ncu-rep: joins the SASS source page with the per-PC instances of the
smsp__pcsamp_warps_issue_stalled_* metrics.  Usage: python tools/ncu_stalls.py REPORT [--top N] [--min S]"""
import argparse
import csv
import io
import re
import subprocess
import sys

REASONS = ["long_scoreboard", "short_scoreboard", "wait", "mio_throttle", "math_pipe_throttle", "barrier", "branch_resolving",
           "no_instructions", "dispatch_stall", "lg_throttle", "tex_throttle", "membar", "sleeping", "drain", "imc_miss",
           "selected", "not_selected", "misc"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    src = list(csv.reader(io.StringIO(run([a.report, "--page", "source", "--csv"]))))[2:]
    sass = {int(r[0], 16): (r[1].strip(), int(r[2]), int(r[5])) for r in src if r and r[0].startswith("0x")}
    metrics = ",".join(f"smsp__pcsamp_warps_issue_stalled_{r}" for r in REASONS)
    raw = list(csv.reader(io.StringIO(run([a.report, "--page", "raw", "--csv", "--print-metric-instances", "details", "--metrics", metrics]))))
    head, vals = raw[0], raw[-1]
    per_pc = {}
    for name, v in zip(head, vals):
        m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", name)
        if not m:
            continue
        for pc, n in re.findall(r"(0x[0-9a-f]+): (\d+)", v):
            if int(n):
                per_pc.setdefault(int(pc, 16), {})[m.group(1)] = int(n)
    tot = sum(s for _, s, _ in sass.values())
    print(f"# {a.report}: {tot} samples")
    rows = sorted(sass.items(), key=lambda kv: -kv[1][1])[: a.top]
    for pc, (ins, samples, execd) in rows:
        why = ", ".join(f"{k}={v}" for k, v in sorted(per_pc.get(pc, {}).items(), key=lambda kv: -kv[1])[:3])
        print(f"{100 * samples / max(tot, 1):5.1f}%  {ins[:64]:<64}  exec={execd:<9} {why}")


if __name__ == "__main__":
    main()
