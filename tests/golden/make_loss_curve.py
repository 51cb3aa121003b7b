"""Generate tests/golden/loss_curve.pt: a 100-step training curve of the REFERENCE's own trainer.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_loss_curve.py [--steps 100]

What runs is the reference, unmodified, through its own entry point for tests
(``/root/reference/tests/train_scripts/train_text_test.py`` -> ``TextTrainer`` -> ``build_foundation_model`` ->
``build_parallelize_model`` (FSDP2, 2 ranks) -> ``veomni_clip_grad_norm`` -> AdamW) on CPU / gloo with the shim documented
in SURVEY.md Appendix B (gloo backend, ``torch.cpu`` memory-stat stubs, a per-sequence SDPA in the flash-attn slot), all ops eager.
Two runs: FSDP2 mixed precision bf16 params / fp32 reduce (the reference's default,
veomni/arguments/arguments_types.py:241-263) and the same with ``param_dtype=float32``, which bounds how much of any
difference is bf16 rounding rather than algorithm.

Captured inside the run by wrapping ``BaseTrainer.forward_backward_step`` / ``on_step_end`` (no reference file is
touched): the initial weights (full tensors, gathered from the DTensor shards at the first step), every micro-batch every
rank sees (input_ids, labels, position_ids, cu_seq_lens) and, per step, the rank's logged loss and the clipped grad-norm.
The global mean loss of a step is the mean over ranks of the logged losses (``mean_global_loss`` scales each rank's
token-weighted loss by ``fsdp_size``, veomni/utils/loss_utils.py:54-90).

tests/test_loss_curve_gpu.py replays the same weights and micro-batches through the sm_100a path on ONE GPU (the two
ranks' micro-batches become two accumulation micro-steps, the same arithmetic mean) and compares the curves.
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = "/root/reference"

TOY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
           head_dim=64, vocab_size=1024, max_position_embeddings=1024, rms_norm_eps=1e-6, tie_word_embeddings=False,
           rope_theta=1000000.0, architectures=["Qwen3ForCausalLM"], model_type="qwen3")
SAMPLE_LEN, MAX_SEQ = 256, 512
LR = 1e-3


def worker():
    """One rank of the reference run (under torchrun)."""
    sys.path.insert(0, REF)
    import torch
    import torch.distributed as dist
    from torch.distributed._tensor import DTensor

    import veomni.utils.device as dev

    dev.get_dist_comm_backend = lambda: "gloo"  # device.py:68-75 raises on CPU
    for n, v in dict(get_device_name=lambda *a: "CPU", max_memory_allocated=lambda *a: 0, max_memory_reserved=lambda *a: 0,
                     memory_stats=lambda *a: {"num_alloc_retries": 0}, memory_allocated=lambda *a: 0,
                     memory_reserved=lambda *a: 0, empty_cache=lambda *a: None,
                     reset_peak_memory_stats=lambda *a: None).items():
        setattr(torch.cpu, n, v)  # EnvironMeter / print_device_mem_info
    dist.init_process_group("gloo")  # the trainer skips init if initialised (trainer/base.py:215-216)
    import veomni.trainer.base as tb

    cap = {"init": None, "micro": [], "steps": []}
    keys = ("input_ids", "labels", "position_ids", "cu_seq_lens_q", "cu_seq_lens_k", "max_length_q", "max_length_k")
    orig_fb = tb.BaseTrainer.forward_backward_step

    def fb(self, micro_batch):
        if cap["init"] is None:
            cap["init"] = {k: (v.full_tensor() if isinstance(v, DTensor) else v).detach().float().clone()
                           for k, v in self.model.state_dict().items()}
        rec = {}
        for k in keys:
            if k in micro_batch:
                v = micro_batch[k]
                rec[k] = v.detach().clone() if isinstance(v, torch.Tensor) else v
        cap["micro"].append((self.state.global_step, rec))
        return orig_fb(self, micro_batch)

    tb.BaseTrainer.forward_backward_step = fb
    # Packed micro-batches: the GPU path of the reference is flash-attn varlen driven by cu_seq_lens (block-diagonal causal
    # attention, veomni/ops/kernels/attention/__init__.py:282-320). flash-attn does not exist on CPU, so the run keeps
    # attn_implementation=flash_attention_2 — VeOmni's own wrapper and kwargs plumbing — and fills the documented slot
    # `_flash_attention_forward` (:33-37) with a per-sequence SDPA of the same semantics. HF's plain
    # "sdpa" path would NOT do: given the collator's all-ones attention_mask it attends across the samples of a pack.
    import torch.nn.functional as F

    import veomni.ops.kernels.attention as A

    def cpu_varlen(query, key, value, attention_mask, query_length=None, is_causal=True, dropout=0.0, softmax_scale=None,
                   cu_seq_lens_q=None, cu_seq_lens_k=None, **kw):  # q/k/v: [B, S, H, D]
        out = torch.empty_like(query)
        Hq, Hk = query.shape[2], key.shape[2]
        cu = cu_seq_lens_q.tolist() if cu_seq_lens_q is not None else [0, query.shape[1]]
        for a, b in zip(cu[:-1], cu[1:]):
            q, k, v = (t[:, a:b].transpose(1, 2) for t in (query, key, value))
            k, v = k.repeat_interleave(Hq // Hk, 1), v.repeat_interleave(Hq // Hk, 1)
            out[:, a:b] = F.scaled_dot_product_attention(q, k, v, is_causal=is_causal, scale=softmax_scale).transpose(1, 2)
        return out

    A._flash_attention_forward = cpu_varlen
    orig_end = tb.BaseTrainer.on_step_end

    def on_end(self, loss=None, loss_dict=None, grad_norm=None):
        gn = grad_norm.item() if hasattr(grad_norm, "item") else float(grad_norm)
        cap["steps"].append((self.state.global_step, float(loss), gn))
        return orig_end(self, loss=loss, loss_dict=loss_dict, grad_norm=grad_norm)

    tb.BaseTrainer.on_step_end = on_end
    out = os.environ["VB_CAPTURE_OUT"]
    orig_destroy = tb.BaseTrainer.destroy_distributed

    def destroy(self):
        torch.save(cap, f"{out}/capture_rank{dist.get_rank()}.pt")
        return orig_destroy(self)

    tb.BaseTrainer.destroy_distributed = destroy
    script = f"{REF}/tests/train_scripts/train_text_test.py"
    sys.argv = [script] + sys.argv[sys.argv.index("--worker") + 1:]
    runpy.run_path(script, run_name="__main__")


def run_reference(tmp: Path, data_dir: Path, cfg_dir: Path, steps: int, mixed: bool) -> list:
    out = tmp / ("bf16" if mixed else "fp32")
    out.mkdir()
    eager = ["attn_implementation=flash_attention_2", "moe_implementation=eager", "cross_entropy_loss_implementation=eager",
             "rms_norm_implementation=eager", "swiglu_mlp_implementation=eager", "rotary_pos_emb_implementation=eager",
             "load_balancing_loss_implementation=eager", "rms_norm_gated_implementation=eager",
             "causal_conv1d_implementation=eager", "chunk_gated_delta_rule_implementation=eager"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc_per_node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(Path(__file__).resolve()), "--worker",
           f"--model.config_path={cfg_dir}", f"--data.train_path={data_dir}", "--data.dyn_bsz_buffer_size=1",
           f"--data.max_seq_len={MAX_SEQ}", "--train.global_batch_size=2", "--train.micro_batch_size=1",
           "--train.init_device=meta", "--train.accelerator.fsdp_config.fsdp_mode=fsdp2", "--train.bsz_warmup_ratio=0",
           "--train.num_train_epochs=1", f"--train.max_steps={steps}", "--train.checkpoint.save_epochs=0",
           "--train.checkpoint.save_steps=0", "--train.checkpoint.save_hf_weights=False",
           f"--train.checkpoint.output_dir={out}", f"--train.optimizer.lr={LR}", "--train.optimizer.weight_decay=0",
           "--train.optimizer.max_grad_norm=1.0", "--train.optimizer.lr_decay_style=constant",
           f"--train.accelerator.fsdp_config.mixed_precision.param_dtype={'bfloat16' if mixed else 'float32'}", "--train.seed=42"]
    cmd += [f"--model.ops_implementation.{e}" for e in eager]
    env = dict(os.environ, VB_CAPTURE_OUT=str(out), OMP_NUM_THREADS="4", PYTHONPATH=f"{REF}:{REF}/tests")
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.exit(res.stdout[-4000:] + "\n" + res.stderr[-6000:])
    import torch

    return [torch.load(out / f"capture_rank{r}.pt", weights_only=False) for r in range(2)], json.loads((out / "log_dict.json").read_text())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    sys.path.insert(0, REF)
    sys.path.insert(0, f"{REF}/tests")
    import torch
    from datasets import Dataset

    from veomni.data.dummy_dataset import build_dummy_dataset

    with tempfile.TemporaryDirectory() as d:
        tmp = Path(d)
        cfg_dir = tmp / "cfg"
        cfg_dir.mkdir()
        (cfg_dir / "config.json").write_text(json.dumps(TOY))
        data_dir = tmp / "data"
        data_dir.mkdir()
        torch.manual_seed(1234)  # DummyTextDataset draws from the global generator (veomni/data/dummy_dataset.py:41-48)
        n = a.steps * 2 * (MAX_SEQ // SAMPLE_LEN) + 64
        ds = build_dummy_dataset("text", n, SAMPLE_LEN)
        rows = [ds[i][0] for i in range(n)]
        # DummyTextDataset's uniform tokens cannot be learned (the curve would sit at ln(1024)); keep its format but give the
        # ids structure — x[t+1] = (5 x[t] + 7) mod 1024 with probability 0.9, a fresh uniform token otherwise — so that the
        # 100 steps trace a real descent with real gradients
        g = torch.Generator().manual_seed(4321)
        for r in rows:
            ids = r["input_ids"].clone()
            fresh = torch.rand(SAMPLE_LEN, generator=g) < 0.1
            for t in range(1, SAMPLE_LEN):
                if not fresh[t]:
                    ids[t] = (5 * ids[t - 1] + 7) % 1024
            r["input_ids"] = ids
            r["labels"] = ids.clone()
            r["labels"][0] = -100
        half = n // 2
        for si, part in enumerate((rows[:half], rows[half:])):
            Dataset.from_list([{k: v.tolist() for k, v in r.items()} for r in part]).to_parquet(str(data_dir / f"{si}.parquet"))
        runs = {}
        for mixed in (True, False):
            caps, log = run_reference(tmp, data_dir, cfg_dir, a.steps, mixed)
            steps = len(caps[0]["steps"])
            loss = [sum(c["steps"][s][1] for c in caps) / len(caps) for s in range(steps)]
            gn = [caps[0]["steps"][s][2] for s in range(steps)]
            assert all(abs(caps[1]["steps"][s][2] - gn[s]) < 1e-5 * max(1, gn[s]) for s in range(steps)), "grad-norm differs across ranks"
            assert [round(x, 6) for x in log["grad_norm"]] == [round(x, 6) for x in gn]
            runs["bf16" if mixed else "fp32"] = {"loss": loss, "grad_norm": gn, "rank_logged_loss": [[c["steps"][s][1] for c in caps] for s in range(steps)]}
            if mixed:
                init = caps[0]["init"]
                micro = []  # [step][rank] -> list of micro-batches
                for s in range(1, steps + 1):
                    micro.append([[m for (gs, m) in c["micro"] if gs == s] for c in caps])
            else:  # both runs must have seen the same data and started from the same weights
                for k, v in caps[0]["init"].items():
                    assert torch.equal(v, init[k]), f"initial weight {k} differs between the bf16 and fp32 runs"
            print(f"reference {'bf16 mixed precision' if mixed else 'fp32'}: loss {loss[0]:.4f} -> {loss[-1]:.4f}, grad_norm {gn[0]:.4f} -> {gn[-1]:.4f}")
        d_loss = max(abs(x - y) for x, y in zip(runs["bf16"]["loss"], runs["fp32"]["loss"]))
        print(f"max |loss_bf16 - loss_fp32| over {a.steps} steps: {d_loss:.3e}")
        # compact: ids < 1024 fit int16, positions int16
        def pack(m):
            out = {}
            for k, v in m.items():
                if isinstance(v, torch.Tensor):
                    out[k] = v.to(torch.int16) if k in ("input_ids", "labels", "position_ids") else v.to(torch.int32)
                else:
                    out[k] = v
            return out

        torch.save({"config": TOY, "init_state_dict": init, "micro_batches": [[[pack(m) for m in r] for r in s] for s in micro],
                    "lr": LR, "betas": (0.9, 0.95), "eps": 1e-8, "weight_decay": 0.0, "max_grad_norm": 1.0,
                    "reference": runs, "world_size": 2,
                    "how": "reference TextTrainer on 2-rank gloo FSDP2 (tests/golden/make_loss_curve.py)"}, HERE / "loss_curve.pt")
        print("wrote", HERE / "loss_curve.pt")


if __name__ == "__main__":
    if "--worker" in sys.argv:
        worker()
    else:
        main()
