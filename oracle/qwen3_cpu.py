"""CPU restatement of one Qwen3 decoder-layer training step (the reference's eager path).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  Used (a) by tests as the model-level oracle
and (b) by ``bench.py`` as the timed CPU baseline (``cpu_baseline`` / ``--impl reference``): the
reference itself is pure Python + PyTorch and cannot travel to the GPU box, so its eager CPU path is
restated here op by op (decoder layer: veomni/models/transformers/qwen3/generated/
patched_modeling_qwen3_gpu.py:347-376; attention :294-333 with eager_attention_forward :236-261).
"""

from __future__ import annotations

import time

import torch

from . import attention as o_attn
from . import ops as o_ops

F = torch.nn.functional


class CPUDecoderLayer(torch.nn.Module):
    def __init__(self, hidden, inter, hq, hk, d, eps=1e-6, dtype=torch.bfloat16, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)

        def w(*shape):
            return torch.nn.Parameter((0.02 * torch.randn(*shape, generator=g)).to(dtype))

        self.hq, self.hk, self.d, self.eps = hq, hk, d, eps
        self.q, self.k, self.v, self.o = w(hq * d, hidden), w(hk * d, hidden), w(hk * d, hidden), w(hidden, hq * d)
        self.gate, self.up, self.down = w(inter, hidden), w(inter, hidden), w(hidden, inter)
        self.ln1 = torch.nn.Parameter(torch.ones(hidden, dtype=dtype))
        self.ln2 = torch.nn.Parameter(torch.ones(hidden, dtype=dtype))
        self.qn = torch.nn.Parameter(torch.ones(d, dtype=dtype))
        self.kn = torch.nn.Parameter(torch.ones(d, dtype=dtype))

    def forward(self, h, cos, sin, seq_lens):
        T = h.shape[0]
        x = o_ops.rms_norm(h, self.ln1, self.eps)
        q = o_ops.rms_norm(F.linear(x, self.q).view(T, self.hq, self.d), self.qn, self.eps)
        k = o_ops.rms_norm(F.linear(x, self.k).view(T, self.hk, self.d), self.kn, self.eps)
        v = F.linear(x, self.v).view(T, self.hk, self.d)
        qe, ke = o_ops.apply_rotary_pos_emb(q.transpose(0, 1)[None], k.transpose(0, 1)[None], cos[None], sin[None])
        outs, off = [], 0
        for n in seq_lens:  # per-sequence causal SDPA == what cu_seqlens tells flash-attn
            qs, ks, vs = qe[:, :, off : off + n], ke[:, :, off : off + n], v[off : off + n].transpose(0, 1)[None]
            outs.append(F.scaled_dot_product_attention(qs, ks, vs, is_causal=True, enable_gqa=True)[0].transpose(0, 1))
            off += n
        a = torch.cat(outs, dim=0).reshape(T, -1)
        h = h + F.linear(a, self.o)
        x = o_ops.rms_norm(h, self.ln2, self.eps)
        return h + F.linear(o_ops.silu_mul(F.linear(x, self.gate), F.linear(x, self.up)), self.down)


def time_layer_step(hidden=4096, inter=12288, hq=32, hk=8, d=128, tokens=512, iters=2, threads=None,
                    dtype=torch.bfloat16):
    """Seconds for one fwd+bwd(+recompute fwd, as with gradient checkpointing) of one layer on `tokens`."""
    if threads:
        torch.set_num_threads(threads)
    layer = CPUDecoderLayer(hidden, inter, hq, hk, d, dtype=dtype)
    cos, sin = o_ops.rotary_cos_sin(torch.arange(tokens)[None], d, 1e6, dtype)
    h = torch.randn(tokens, hidden).to(dtype).requires_grad_(True)
    best = float("inf")
    for _ in range(iters):
        t0 = time.perf_counter()
        with torch.no_grad():
            layer(h, cos[0], sin[0], [tokens])  # checkpointed forward
        out = layer(h, cos[0], sin[0], [tokens])  # recompute
        out.float().square().mean().backward()
        best = min(best, time.perf_counter() - t0)
    return best


class CPUSampleModel(torch.nn.Module):
    """The full Qwen3 causal LM of BASELINE configs[1] for *timing*: embedding, ``layers`` decoder layers with
    per-layer gradient checkpointing (veomni/distributed/torch_parallelize.py:445-456), final norm, lm_head and the
    causal-LM loss (veomni/ops/kernels/cross_entropy/__init__.py:180-221, eager path: bf16 logits upcast to fp32).
    To stay inside a bounded amount of host memory the decoder layers share ONE set of weights (each layer still
    streams its 386 MB of bf16 weights from DRAM — nothing that large stays in cache); embedding and lm_head are the
    full [151936, 4096] matrices."""

    def __init__(self, layers=36, hidden=4096, inter=12288, hq=32, hk=8, d=128, vocab=151936, dtype=torch.bfloat16):
        super().__init__()
        self.layers, self.d = layers, d
        self.layer = CPUDecoderLayer(hidden, inter, hq, hk, d, dtype=dtype)
        g = torch.Generator().manual_seed(1)

        def big(rows, cols):  # timing only: a 4 M-element random block tiled to size (initialisation is not the workload)
            base = (0.02 * torch.randn(1 << 22, generator=g)).to(dtype)
            return torch.nn.Parameter(base.repeat(rows * cols // base.numel() + 1)[: rows * cols].view(rows, cols).clone())

        self.embed, self.lm_head = big(vocab, hidden), big(vocab, hidden)
        self.norm = torch.nn.Parameter(torch.ones(hidden, dtype=dtype))

    def forward(self, ids, labels, cos, sin):
        from torch.utils.checkpoint import checkpoint

        h = F.embedding(ids, self.embed)
        for _ in range(self.layers):
            h = checkpoint(self.layer, h, cos, sin, [ids.numel()], use_reentrant=False)
        h = o_ops.rms_norm(h, self.norm, self.layer.eps)
        logits = F.linear(h, self.lm_head).float()
        shift = F.pad(labels, (0, 1), value=-100)[1:]
        return F.cross_entropy(logits, shift, ignore_index=-100)


class SampleStep:
    """One bounded sample of the Qwen3-8B training step on host cores: the whole model (36 layers, embedding, lm_head,
    loss) forward + recompute + backward on ``tokens`` tokens, gradient-norm clip, and an AdamW step over the same
    fraction ``tokens / 4096`` of the 8.19 B fp32 master parameters (so the sample keeps the step's ratio of per-token
    work to per-step optimizer work). ``run()`` returns the seconds of one such sample step."""

    def __init__(self, tokens=256, layers=36, threads=None, seq_len=4096, n_params=8_190_735_360):
        if threads:
            torch.set_num_threads(threads)
        self.tokens = tokens
        self.model = CPUSampleModel(layers=layers)
        g = torch.Generator().manual_seed(2)
        self.ids = torch.randint(0, 1024, (tokens,), generator=g)
        self.labels = self.ids.clone()
        self.labels[0] = -100
        cos, sin = o_ops.rotary_cos_sin(torch.arange(tokens)[None], self.model.d, 1e6, torch.bfloat16)
        self.cos, self.sin = cos[0], sin[0]
        n_opt = int(n_params * tokens / seq_len)
        self.master = torch.nn.Parameter(torch.zeros(n_opt))
        self.master.grad = torch.zeros(n_opt)
        self.opt = torch.optim.AdamW([self.master], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0)

    def run(self) -> float:
        t0 = time.perf_counter()
        for p in self.model.parameters():
            p.grad = None
        loss = self.model(self.ids, self.labels, self.cos, self.sin)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(self.model.parameters()), 1.0)
        self.opt.step()
        self.loss = float(loss.detach())
        return time.perf_counter() - t0
