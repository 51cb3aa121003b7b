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
