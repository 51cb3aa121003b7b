"""Top SASS instructions by warp-stall samples from an .ncu-rep source page, plus an opcode histogram.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [N]
"""
import collections
import csv
import subprocess
import sys


def main(path, n=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    S, I, X = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= X or not r[I].isdigit():
            if r and r[0] == "Kernel Name":
                break
            continue
        data.append((int(r[I]), int(r[X] or 0), r[S].strip()))
    tot = sum(d[0] for d in data) or 1
    print(f"# {path}: {len(data)} SASS instructions, {tot} stall samples")
    print("top instructions by samples:")
    for s, x, src in sorted(data, reverse=True)[:n]:
        print(f"  {100*s/tot:5.1f}%  exec={x:>9d}  {src}")
    ops = collections.Counter()
    execs = collections.Counter()
    for s, x, src in data:
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        ops[op] += s
        execs[op] += x
    print("by opcode (samples %, warp-instructions executed):")
    for op, s in ops.most_common(18):
        print(f"  {100*s/tot:5.1f}%  {execs[op]:>11d}  {op}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
