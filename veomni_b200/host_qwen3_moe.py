"""Qwen3-MoE caller of the hot path: the dense Qwen3 caller with the MLP replaced by a sparse MoE block.

Mirrors veomni/models/transformers/qwen3_moe/generated/patched_modeling_qwen3_moe_gpu.py:
``Qwen3MoeTopKRouter`` (:312-336: linear -> fp32 softmax -> top-k -> renormalise -> cast),
``Qwen3MoeExperts`` (:255-300: merged ``gate_up_proj [E, 2I, H]``, ``down_proj [E, H, I]``, OpSlot
``moe_experts``) and ``Qwen3MoeSparseMoeBlock`` (:339-377).  Expert computation goes through
``veomni_b200.moe.moe_experts_forward`` (routing kernels + tcgen05 GroupGEMM; EP dispatch/combine over NVLink
when an ``EPContext`` is installed with ``veomni_b200.moe.set_ep_group``).
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as Fnn

from . import moe
from .host_qwen3 import Qwen3Attention, Qwen3Config, Qwen3ForCausalLM, Qwen3RMSNorm


@dataclass
class Qwen3MoeConfig(Qwen3Config):
    num_experts: int = 128
    num_experts_per_tok: int = 8
    moe_intermediate_size: int = 768
    norm_topk_prob: bool = True

    @staticmethod
    def qwen3_30b_a3b() -> "Qwen3MoeConfig":
        return Qwen3MoeConfig(hidden_size=2048, intermediate_size=6144, num_hidden_layers=48, num_attention_heads=32,
                              num_key_value_heads=4, head_dim=128, num_experts=128, num_experts_per_tok=8,
                              moe_intermediate_size=768)


class Qwen3MoeTopKRouter(nn.Module):
    def __init__(self, cfg: Qwen3MoeConfig):
        super().__init__()
        self.top_k, self.num_experts, self.norm_topk_prob = cfg.num_experts_per_tok, cfg.num_experts, cfg.norm_topk_prob
        self.weight = nn.Parameter(torch.zeros(cfg.num_experts, cfg.hidden_size))

    def forward(self, x):
        logits = Fnn.linear(x, self.weight)
        probs = Fnn.softmax(logits, dtype=torch.float, dim=-1)
        top_v, top_i = torch.topk(probs, self.top_k, dim=-1)
        if self.norm_topk_prob:
            top_v = top_v / top_v.sum(dim=-1, keepdim=True)
        return logits, top_v.to(logits.dtype), top_i


class Qwen3MoeExperts(nn.Module):
    def __init__(self, cfg: Qwen3MoeConfig):
        super().__init__()
        self.num_experts = cfg.num_experts
        self.gate_up_proj = nn.Parameter(torch.empty(cfg.num_experts, 2 * cfg.moe_intermediate_size, cfg.hidden_size))
        self.down_proj = nn.Parameter(torch.empty(cfg.num_experts, cfg.hidden_size, cfg.moe_intermediate_size))

    def forward(self, hidden_states, top_k_index, top_k_weights):  # OpSlot("moe_experts", "standard") call site
        return moe.moe_experts_forward(self, hidden_states, top_k_index, top_k_weights)


class Qwen3MoeSparseMoeBlock(nn.Module):
    def __init__(self, cfg: Qwen3MoeConfig):
        super().__init__()
        self.experts = Qwen3MoeExperts(cfg)
        self.gate = Qwen3MoeTopKRouter(cfg)

    def forward(self, x):
        _, w, idx = self.gate(x)
        return self.experts(x, idx, w)


class Qwen3MoeDecoderLayer(nn.Module):
    def __init__(self, cfg: Qwen3MoeConfig, layer_idx: int):
        super().__init__()
        self.self_attn = Qwen3Attention(cfg, layer_idx)
        self.mlp = Qwen3MoeSparseMoeBlock(cfg)
        self.input_layernorm = Qwen3RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = Qwen3RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)

    def forward(self, h, cos, sin, cu_seqlens, max_seqlen, sp_group=None, fwd_id=None):
        h = h + self.self_attn(self.input_layernorm(h), cos, sin, cu_seqlens, max_seqlen, sp_group, fwd_id)
        return h + self.mlp(self.post_attention_layernorm(h))


class Qwen3MoeForCausalLM(Qwen3ForCausalLM):
    _no_split_modules = ["Qwen3MoeDecoderLayer"]

    def __init__(self, cfg: Qwen3MoeConfig):
        super().__init__(cfg)
        self.model.layers = nn.ModuleList([Qwen3MoeDecoderLayer(cfg, i) for i in range(cfg.num_hidden_layers)])

    @torch.no_grad()
    def init_weights(self, seed: int = 0):
        from .host_qwen3 import _is_dtensor, _local, _rank_seed

        super().init_weights(seed)
        sharded = _is_dtensor(self.lm_head.weight)
        g = torch.Generator(device=self.lm_head.weight.device).manual_seed(seed + 1 + (_rank_seed() if sharded else 0))
        for m in self.modules():
            if isinstance(m, Qwen3MoeExperts):
                _local(m.gate_up_proj).normal_(0.0, self.config.initializer_range, generator=g)
                _local(m.down_proj).normal_(0.0, self.config.initializer_range, generator=g)
            elif isinstance(m, Qwen3MoeTopKRouter):
                _local(m.weight).normal_(0.0, self.config.initializer_range, generator=g)

    def get_parallel_plan(self):
        from .parallel_plan import qwen3_moe_parallel_plan

        return qwen3_moe_parallel_plan()
