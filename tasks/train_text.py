"""``python tasks/train_text.py --model.config_path <dir-or-qwen3_8b> --train.max_steps N ...``

Entry point with the same name and role as the reference's ``tasks/train_text.py`` (tasks/train_text.py:1-8 ->
``TextTrainer``; loop: veomni/trainer/text_trainer.py:103-183), restricted to the hot path this repository
implements: Qwen3 dense text model, FSDP2 (+Ulysses), synthetic ``DummyTextDataset``-style packed batches
(veomni/data/dummy_dataset.py:26-48), AdamW, grad-norm clipping, per-step loss / grad_norm / tokens-per-second
log lines.  With the reference installed, keep using ITS entry point and select the ``b200`` implementations in
the YAML; this file exists so the path is runnable where the reference is not installed.
Launch with torchrun for more than one GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model.config_path", dest="config_path", default="qwen3_8b")
    ap.add_argument("--data.max_seq_len", dest="max_seq_len", type=int, default=4096)
    ap.add_argument("--data.samples_per_pack", dest="samples_per_pack", type=int, default=1)
    ap.add_argument("--train.max_steps", dest="max_steps", type=int, default=10)
    ap.add_argument("--train.lr", dest="lr", type=float, default=1e-4)
    ap.add_argument("--train.max_grad_norm", dest="max_grad_norm", type=float, default=1.0)
    ap.add_argument("--train.seed", dest="seed", type=int, default=42)
    ap.add_argument("--train.accelerator.ulysses_size", dest="ulysses_size", type=int, default=1)
    ap.add_argument("--train.enable_gradient_checkpointing", dest="ckpt", type=int, default=1)
    ap.add_argument("--train.output_dir", dest="output_dir", default="")
    ap.add_argument("--model.num_hidden_layers", dest="layers", type=int, default=0, help="override (debug / smoke runs)")
    return ap.parse_args()


def main():
    a = parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29519")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    from veomni_b200.clip_grad_norm import clip_grad_norm
    from veomni_b200.host_qwen3 import Qwen3Config, Qwen3ForCausalLM
    from veomni_b200.parallel_state import init_parallel_state
    from veomni_b200.parallelize import build_parallelize_model

    ps = init_parallel_state(dp_size=world // a.ulysses_size, ulysses_size=a.ulysses_size)
    if a.config_path == "qwen3_8b":
        cfg = Qwen3Config.qwen3_8b()
    else:
        cfg = Qwen3Config.from_hf_dict(json.loads((Path(a.config_path) / "config.json").read_text()))
    if a.layers:
        cfg.num_hidden_layers = a.layers
    with torch.device("meta"):
        model = Qwen3ForCausalLM(cfg)
    model.to_empty(device=dev)
    model.inv_freq.copy_(1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, device=dev).float() / cfg.head_dim)))
    model.init_weights(seed=0)
    model = build_parallelize_model(model, enable_gradient_checkpointing=bool(a.ckpt), mesh=ps.fsdp_mesh)
    model.sp_group = ps.ulysses_group
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=a.lr, betas=(0.9, 0.95), fused=True)

    P, sp_rank = a.ulysses_size, ps.ulysses_rank
    dp_rank = rank // P
    g = torch.Generator().manual_seed(a.seed + dp_rank)
    L = a.max_seq_len // a.samples_per_pack
    log = []
    for step in range(a.max_steps):
        ids = torch.randint(0, 1024, (1, a.max_seq_len), generator=g)
        labels = ids.clone()
        labels[0, ::L] = -100
        pos = torch.arange(L).repeat(a.samples_per_pack)[None]
        cu = torch.arange(0, a.max_seq_len + 1, L, dtype=torch.int32, device=dev)
        shift = None
        if P > 1:  # SequenceParallelCollator (data_collator.py:336-389): shift labels, then slice the row
            shift = torch.nn.functional.pad(labels, (0, 1), value=-100)[..., 1:]
            sl = slice(sp_rank * a.max_seq_len // P, (sp_rank + 1) * a.max_seq_len // P)
            ids, pos, shift = ids[:, sl], pos[:, sl], shift[:, sl].to(dev)
        t0 = time.perf_counter()
        loss = model(ids.to(dev), pos.to(dev), cu, L, labels=None if P > 1 else labels.to(dev), shift_labels=shift)
        if P > 1:  # reduce_sequence_parallel_loss (sequence_parallel/loss.py:27-64): token-weighted mean over SP ranks
            n = (shift != -100).sum().float()
            num, den = loss * n, n.clone()
            dist.all_reduce(den, group=ps.ulysses_group)
            loss = num / den * P  # gradients are averaged over the FSDP group that includes the SP ranks
        loss.backward()
        gn = clip_grad_norm(model, a.max_grad_norm)
        opt.step()
        opt.zero_grad(set_to_none=True)
        lv = loss.detach().float()
        if P > 1:
            lv = lv / P
            dist.all_reduce(lv, group=ps.ulysses_group)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rec = {"step": step, "loss": round(float(lv), 6), "grad_norm": round(float(gn), 6),
               "tokens_per_second": round(a.max_seq_len * (world // P) / dt, 1)}
        log.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if rank == 0 and a.output_dir:
        Path(a.output_dir).mkdir(parents=True, exist_ok=True)
        (Path(a.output_dir) / "log_dict.json").write_text(json.dumps(log))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
