"""Build and run tools/ubench_sm100.cu (pipe-throughput microbenchmarks for the attention softmax stage).

    python tools/ubench.py            # on a B200: prints one JSON line per measurement
    python tools/ubench.py --build    # compile only (works without a GPU)
"""
import argparse
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE.parent / "build" / "ubench_sm100"  # build/ is git-ignored


def build() -> Path:
    OUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--use_fast_math",
           "-o", str(OUT), str(HERE / "ubench_sm100.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        sys.exit(res.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    a = ap.parse_args()
    exe = build()
    if not a.build:
        sys.exit(subprocess.run([str(exe)]).returncode)
