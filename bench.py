#!/usr/bin/env python
"""bench.py — tokens/sec of the Qwen3-8B FSDP2 bf16 training step (seq_len 4096, synthetic packed text).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's eager CPU path on this box's host cores

One "step" = forward + backward (per-layer gradient checkpointing, as VeOmni's default) + grad-norm clip +
fused AdamW step + zero_grad of the full Qwen3-8B (36 layers, 8.19 B params, random init) under FSDP2
(bf16 params / fp32 reduce), one packed 4096-token micro-batch per rank (weak scaling).  Every op on the
hot path (RMSNorm, q/k-norm+RoPE, varlen attention, SwiGLU, FSDP2 all-gather / reduce-scatter) is a
veomni_b200 sm_100a kernel; dense projections are cuBLAS.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

SEQ_LEN = 4096
METRIC = "tokens/sec (Qwen3-8B FSDP2 bf16 seq4096)"


# DRAM bytes per launch (read + write) of the three attention kernels at T=4096, 32/8 heads, D=128, from ncu --set full
NCU_TRAFFIC_BYTES = {"bwd_dkdv": 86041088 + 4705024, "bwd_dq": 84966656 + 11869440, "fwd": 50374400 + 3454208}


def _peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "how": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "how": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_baseline(sample_tokens: int, iters: int, threads: int | None) -> dict:
    """The reference's eager CPU path (restated in oracle/qwen3_cpu.py) on a bounded sample."""
    import torch

    from oracle.qwen3_cpu import time_layer_step

    ncpu = os.cpu_count() or 1
    if threads:
        cores = threads
    else:
        # all host threads are available to the reference arm, but small bf16 GEMMs stop scaling (and slow down) long
        # before 128 threads: pick the fastest thread count on a 64-token probe and report the count actually used
        cand = sorted({c for c in (ncpu, ncpu // 2, ncpu // 4, 32, 16) if 1 <= c <= ncpu})
        probe = {c: time_layer_step(tokens=64, iters=1, threads=c) for c in cand}
        cores = min(probe, key=probe.get)
    torch.set_num_threads(cores)
    t_layer = time_layer_step(tokens=sample_tokens, iters=iters, threads=cores)
    # 36 identical layers; per-token cost measured at `sample_tokens` (attention's quadratic term is
    # under-counted at 4096; embedding, lm_head, loss and optimizer are not counted: an optimistic CPU number)
    step_s = t_layer * 36 * (SEQ_LEN / sample_tokens)
    return {"value": round(SEQ_LEN / step_s, 3), "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"1 of 36 Qwen3-8B decoder layers, fwd + recompute + bwd in bf16 on {sample_tokens} tokens, "
                      f"best of {iters}; scaled x36 layers x{SEQ_LEN // sample_tokens} to a 4096-token step",
            "seconds_per_sample": round(t_layer, 3)}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    base = cpu_baseline(sample_tokens=512, iters=max(1, min(args.steps, 3)), threads=None)
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(SEQ_LEN / base["value"] * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "Qwen3-8B FSDP2 bf16 seq_len 4096 training step (BASELINE configs[1]), reference eager ops on host CPU",
                      "global_batch": 1, "seq_len": SEQ_LEN},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (see module docstring)")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    from veomni_b200 import _lib
    from veomni_b200 import attention as vattn
    from veomni_b200.clip_grad_norm import clip_grad_norm
    from veomni_b200.host_qwen3 import Qwen3Config, Qwen3ForCausalLM, flops_per_token
    from veomni_b200.parallelize import build_parallelize_model

    _lib.load()  # fail loudly if the CUDA library is missing
    cfg = Qwen3Config.qwen3_8b()
    if args.layers:
        cfg.num_hidden_layers = args.layers  # debugging only; reported in config and marks the run invalid
    with torch.device("meta"):
        model = Qwen3ForCausalLM(cfg)
    model.to_empty(device=dev)
    model.inv_freq.copy_(1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, device=dev).float() / cfg.head_dim)))
    model.init_weights(seed=0)
    model = build_parallelize_model(model, b200_comm=not args.nccl_comm, comm_ctas=args.comm_ctas)
    model.train()
    if args.b200_adamw:  # experimental multi-tensor AdamW kernel (veomni_b200/optim.py); marks the line
        from veomni_b200.optim import B200AdamW

        opt = B200AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0)
    else:
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0, fused=True)

    # synthetic packed micro-batches in pinned host memory (DummyTextDataset: ids ~ U{0..1023}, first label ignored)
    g = torch.Generator().manual_seed(1234 + rank)
    nbuf = 4
    host = []
    for _ in range(nbuf):
        ids = torch.randint(0, 1024, (1, SEQ_LEN), generator=g, dtype=torch.int64)
        labels = ids.clone()
        labels[0, 0] = -100
        pos = torch.arange(SEQ_LEN, dtype=torch.int64)[None]
        host.append(tuple(t.pin_memory() for t in (ids, labels, pos)))
    cu = torch.tensor([0, SEQ_LEN], dtype=torch.int32, device=dev)
    resident = [tuple(t.to(dev) for t in h) for h in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def step(batch):
        ids, labels, pos = batch
        loss = model(ids, pos, cu, SEQ_LEN, labels=labels)
        loss.backward()
        clip_grad_norm(model, 1.0)  # veomni_clip_grad_norm semantics, multi-tensor kernels
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(nsteps, e2e: bool):
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        last = None
        for i in range(nsteps):
            if e2e:
                batch = tuple(t.to(dev, non_blocking=True) for t in host[i % nbuf])
                last = step(batch).item()  # device->host read of the step's loss
            else:
                last = step(resident[i % nbuf])
        e.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([s.elapsed_time(e)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / nsteps, (last if isinstance(last, float) else float(last.item()))

    for i in range(args.warmup):
        step(resident[i % nbuf])
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    # CUDA-event pairs around every launch of the three attention kernels (the heaviest kernels of ours)
    vattn.PROFILE = {"fwd": [], "bwd_dq": [], "bwd_dkdv": []}
    ms_dev, loss_dev = timed(args.steps, e2e=False)
    launches = _lib.launch_count()
    prof = vattn.PROFILE
    vattn.PROFILE = None
    torch.cuda.synchronize()
    prof_ms = {k: [a.elapsed_time(b) for a, b in v] for k, v in prof.items()}
    ms_e2e, loss_e2e = timed(args.steps, e2e=True)
    mem_gb = torch.cuda.max_memory_allocated() / 2**30
    # The same step with every activation kept instead of recomputed (180 GB of HBM3e hold them at this size): reported
    # next to the headline, which keeps the reference's default of per-layer gradient checkpointing.
    no_rc = None
    if not args.skip_no_recompute:
        inner = getattr(model, "module", model)
        prev = inner.gradient_checkpointing
        inner.gradient_checkpointing = False
        try:
            torch.cuda.reset_peak_memory_stats()
            for i in range(2):
                step(resident[i % nbuf])
            ms_nr, _ = timed(args.steps, e2e=False)
            no_rc = {"value": round(SEQ_LEN * world / (ms_nr / 1e3), 1), "unit": "tokens/s", "ms_per_step": round(ms_nr, 2),
                     "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
        except torch.OutOfMemoryError:
            no_rc = {"value": None, "note": "out of memory without recomputation"}
        except Exception as ex:  # noqa: BLE001  (a side measurement must never take the headline line down with it)
            no_rc = {"value": None, "note": f"{type(ex).__name__}: {str(ex)[:160]}"}
        inner.gradient_checkpointing = prev
    if args.torch_profile:  # debugging aid, outside every timed region: per-kernel device time of one step on rank 0
        from torch.profiler import ProfilerActivity, profile

        dist.barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            step(resident[0])
            torch.cuda.synchronize()
        if rank == 0:
            with open(args.torch_profile, "w") as fh:
                fh.write(tp.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else {}

    if rank == 0:
        peaks = _peaks()
        tokens = SEQ_LEN * world
        value = tokens / (ms_dev / 1e3)
        fpt = flops_per_token(cfg, [SEQ_LEN])
        fwd_flops = 4 * SEQ_LEN * SEQ_LEN * cfg.head_dim * cfg.num_attention_heads / 2  # causal fwd per launch
        # algorithmic FLOPs per launch: fwd = 2 matmuls, dQ = 1, dK+dV = 2 of the 5 backward matmuls
        kinfo = {"fwd": ("attn_fwd_tc_kernel (varlen causal attention forward, tcgen05)", fwd_flops),
                 "bwd_dkdv": ("attn_bwd_dkdv_tc_kernel (attention backward dK/dV, tcgen05)", fwd_flops * 2.5 * 0.6),
                 "bwd_dq": ("attn_bwd_dq_tc_kernel (attention backward dQ, tcgen05)", fwd_flops * 2.5 * 0.4)}
        totals = {k: sum(v) for k, v in prof_ms.items() if v}
        dom = max(totals, key=totals.get) if totals else "fwd"
        attn_ms = prof_ms.get(dom, [])
        attn_flops = kinfo[dom][1]
        attn_avg_ms = sum(attn_ms) / max(1, len(attn_ms))
        achieved = attn_flops / (attn_avg_ms * 1e-3) / 1e12 if attn_ms else None
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_dev, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Qwen3-8B FSDP2 bf16 seq_len 4096 training step (BASELINE configs[1]): fwd+bwd with per-layer "
                                   "gradient checkpointing, clip_grad_norm, fused AdamW; 1x4096-token packed sample per rank",
                       "global_batch": world, "seq_len": SEQ_LEN, "parallelism": f"fsdp{world}", "layers": cfg.num_hidden_layers,
                       "params_b": round(cfg.num_params() / 1e9, 3), "l2": "inputs (16 GB of bf16 weights per step) larger than L2",
                       "fsdp_comm": "nccl" if args.nccl_comm else "veomni_b200 NVLink pull kernels",
                       "optimizer": "veomni_b200 multi-tensor AdamW (experimental)" if args.b200_adamw else "torch.optim.AdamW(fused=True)",
                       "valid": args.layers == 0},
            "tokens_per_sec_per_gpu": round(value / world, 1),
            "mfu_measured_peak": round(fpt * value / world / (peaks["bf16_tflops_sustained"] * 1e12), 4),
            "mfu_2250": round(fpt * value / world / 2250e12, 4),
            "e2e": {"value": round(tokens / (ms_e2e / 1e3), 1), "unit": "tokens/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e, 2)},
            "gpu_launches": int(launches),
            "roofline": {"kernel": kinfo[dom][0] + f", {len(attn_ms) // max(1, args.steps)} launches/step",
                         "attention_ms_per_step": {k: round(v / max(1, args.steps), 2) for k, v in totals.items()},
                         "bound": "tensor", "achieved": round(achieved, 1) if achieved else None,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": round(achieved / peaks["bf16_tflops_sustained"], 4) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of
                         # the same kernel at the same shape (profiles/r01_attn_bwd_*_tc_ts_ncu.txt, r01_attn_fwd_tc_v3_ncu.txt)
                         "traffic": NCU_TRAFFIC_BYTES.get(dom), "traffic_source": "profiles/r01_attn_*_ncu.txt (ncu --set full)",
                         "peak_source": f"{peaks['how']} (sustained cuBLAS bf16, kernel timed inside a long step)",
                         "avg_launch_ms": round(attn_avg_ms, 4), "launches_timed": len(attn_ms),
                         "algorithmic_flops_per_launch": attn_flops},
            "clocks": clocks, "loss": round(loss_dev, 4), "loss_e2e": round(loss_e2e, 4), "max_mem_gb": round(mem_gb, 1),
        }
        if no_rc is not None:
            out["no_recompute"] = no_rc
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(sample_tokens=256, iters=1, threads=None)
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=0, help="debug only: fewer layers (marks the line invalid)")
    ap.add_argument("--nccl-comm", action="store_true", help="debug: PyTorch's default NCCL FSDP2 collectives")
    ap.add_argument("--comm-ctas", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--b200-adamw", action="store_true", help="experimental: veomni_b200.optim.B200AdamW instead of torch's fused AdamW")
    ap.add_argument("--skip-no-recompute", action="store_true", help="skip the extra no-recomputation measurement")
    ap.add_argument("--torch-profile", default="", help="debug: write a torch.profiler kernel table of one extra step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
