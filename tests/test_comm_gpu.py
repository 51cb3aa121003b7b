"""Multi-GPU parity for the NVLink collectives: spawns tests/multigpu_worker.py under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("nproc", [2])
def test_p2p_collectives_multigpu(nproc):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, PYTHONPATH=str(REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(REPO / "tests" / "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("WORKER OK") == nproc


def test_fsdp_pack_bf16_equals_chunk_cat(cuda_dev):
    """The reduce-scatter copy-in kernel reproduces torch._chunk_cat's layout (bit-exact, bf16 kept), including
    zero padding of ragged dim-0 sizes, the scalar (unaligned) variant and > 24 parameters per call."""
    import ctypes

    from veomni_b200 import _lib
    from veomni_b200.fsdp_comm import pack_plan

    lib = _lib.load()
    g = torch.Generator(device=cuda_dev).manual_seed(11)
    cases = [
        [(4096, 512), (4096,), (1024, 128), (128,), (512, 4096)],          # aligned: vector path
        [(10, 4), (3,), (7, 2, 2), (16, 8), (1, 5)],                        # ragged: scalar path
        [(64, 8)] * 30,                                                     # more parameters than one launch holds
    ]
    for shapes in cases:
        for world in (2, 4, 8):
            grads = [torch.randn(*s, generator=g, device=cuda_dev).to(torch.bfloat16) for s in shapes]
            plan, row = pack_plan(shapes, world)
            ref = torch.empty(world, row, dtype=torch.bfloat16, device=cuda_dev)
            torch._chunk_cat(grads, dim=0, num_chunks=world, out=ref)
            out = torch.full((world, row), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
            flat = []
            for t, (numel, chunk, off) in zip(grads, plan):
                flat += [t.data_ptr(), numel, chunk, off]
            arr = (ctypes.c_int64 * len(flat))(*flat)
            _lib.check(lib.vb200_fsdp_pack_bf16(arr, len(plan), world, row, out.data_ptr(), 0, _lib.stream_ptr()), "pack")
            torch.cuda.synchronize()
            assert torch.equal(out, ref), (shapes[:2], world)
            # fp32 sources (the all-gather copy-in with world = 1 uses the same kernel): rounded once to bf16
            g32 = [t.float() + 1e-3 for t in grads]
            ref32 = torch.empty(world, row, dtype=torch.bfloat16, device=cuda_dev)
            torch._chunk_cat([t.to(torch.bfloat16) for t in g32], dim=0, num_chunks=world, out=ref32)
            flat = []
            for t, (numel, chunk, off) in zip(g32, plan):
                flat += [t.data_ptr(), numel, chunk, off]
            arr = (ctypes.c_int64 * len(flat))(*flat)
            out.fill_(float("nan"))
            _lib.check(lib.vb200_fsdp_pack_bf16(arr, len(plan), world, row, out.data_ptr(), 1, _lib.stream_ptr()), "pack")
            torch.cuda.synchronize()
            assert torch.equal(out, ref32), ("fp32 source", shapes[:2], world)
