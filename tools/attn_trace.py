"""Timeline of the attention backward dQ kernel's hand-offs (block 0, the heaviest Q tile): clock64 stamps recorded inside
the kernel (attention.BWD_TRACE) -> per-tile deltas. Usage: python tools/attn_trace.py [--p16]"""
import argparse
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from veomni_b200 import _lib  # noqa: E402
from veomni_b200 import attention as A  # noqa: E402

EV = ["mma: S/dP_j issued", "mma: dS_j seen, dQ_j issued", "softmax w2: S/dP_j ready", "softmax w2: dS buffer free",
      "softmax w2: chunk in registers", "softmax w2: dS stored", "softmax w2: arrived"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p16", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    T, Hq, Hk, D = 4096, 32, 8, 128
    q, k, v = (torch.randn(T, h, D, device=dev, dtype=torch.bfloat16).requires_grad_(True) for h in (Hq, Hk, Hk))
    cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
    do = torch.randn(T, Hq, D, device=dev, dtype=torch.bfloat16)
    A.FWD_IMPL = A.BWD_IMPL = "tc"
    A.BWD_P16 = a.p16
    for _ in range(2):
        A.flash_attn_varlen(q, k, v, cu, T).backward(do)
    A.BWD_TRACE = True
    A.flash_attn_varlen(q, k, v, cu, T).backward(do)
    A.BWD_TRACE = False
    buf = (ctypes.c_int64 * 512)()
    _lib.check(_lib.load().vb200_attn_debug_trace(buf), "trace")
    t = torch.tensor(list(buf)).view(8, 64)
    t0 = int(t[0, 0])
    rel = (t - t0).tolist()
    print(json.dumps({"p16": a.p16, "events": EV}))
    for j in range(0, 64):
        print(j, [rel[e][j] if t[e, j] else None for e in range(7)])
    # steady-state per-tile period of each event (tiles 8..56)
    per = {EV[e]: round(float(t[e, 56] - t[e, 8]) / 48, 1) for e in range(7) if t[e, 56] and t[e, 8]}
    print(json.dumps({"clk_per_tile": per}))


if __name__ == "__main__":
    main()
