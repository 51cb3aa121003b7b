"""A compact Qwen3 causal-LM *caller* of the hot path.

This is the host-side mirror of the call sites in VeOmni's patched Qwen3 modeling
(veomni/models/transformers/qwen3/generated/patched_modeling_qwen3_gpu.py): the same module tree and
parameter names as HF ``Qwen3ForCausalLM`` (so reference / HF state dicts load unchanged), the same
order of operations in the decoder layer (:347-376) and attention (:294-333), packed ("padding-free")
inputs with ``cu_seq_lens`` as ``MainCollator`` produces them (veomni/data/data_collator.py:392-459),
per-layer gradient checkpointing (GradientCheckpointingLayer), and the causal-LM loss of
``ForCausalLMLoss`` (veomni/ops/kernels/cross_entropy/__init__.py:180-221).

Every op on the path is one of the sm_100a kernels of this package (RMSNorm, fused q/k-norm + RoPE,
varlen attention, SwiGLU, Ulysses all-to-all); the dense projections are ``F.linear`` (cuBLAS — out of
scope).  It exists so that bench.py, smoke() and the parity tests can drive the path
on a box where the reference is not installed; with the reference installed, ``veomni_b200.registry``
plugs the same kernels into its OpSlots instead.
"""

from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as Fnn
from torch.utils.checkpoint import checkpoint

from . import functional as F
from .attention import flash_attn_varlen
from .cross_entropy import b200_cross_entropy


@dataclass
class Qwen3Config:
    vocab_size: int = 151936
    hidden_size: int = 4096
    intermediate_size: int = 12288
    num_hidden_layers: int = 36
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 128
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    tie_word_embeddings: bool = False
    initializer_range: float = 0.02

    @staticmethod
    def qwen3_8b() -> "Qwen3Config":
        return Qwen3Config()

    @staticmethod
    def from_hf_dict(d: dict) -> "Qwen3Config":
        rope = d.get("rope_theta")
        if rope is None and isinstance(d.get("rope_parameters"), dict):
            rope = d["rope_parameters"].get("rope_theta")
        return Qwen3Config(
            vocab_size=d["vocab_size"], hidden_size=d["hidden_size"], intermediate_size=d["intermediate_size"],
            num_hidden_layers=d["num_hidden_layers"], num_attention_heads=d["num_attention_heads"],
            num_key_value_heads=d["num_key_value_heads"],
            head_dim=d.get("head_dim") or d["hidden_size"] // d["num_attention_heads"],
            rms_norm_eps=d.get("rms_norm_eps", 1e-6), rope_theta=rope or 1000000.0,
            tie_word_embeddings=d.get("tie_word_embeddings", False),
        )

    def num_params(self) -> int:
        h, i, L = self.hidden_size, self.intermediate_size, self.num_hidden_layers
        qd, kd = self.num_attention_heads * self.head_dim, self.num_key_value_heads * self.head_dim
        layer = h * qd + 2 * h * kd + qd * h + 3 * h * i + 2 * h + 2 * self.head_dim
        return L * layer + self.vocab_size * h * (1 if self.tie_word_embeddings else 2) + h


class Qwen3RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.variance_epsilon = eps

    def forward(self, x):  # OpSlot("rms_norm", "standard") call site (:90-91)
        return F.rms_norm(x, self.weight, self.variance_epsilon)


class Qwen3MLP(nn.Module):
    def __init__(self, cfg: Qwen3Config):
        super().__init__()
        self.gate_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.up_proj = nn.Linear(cfg.hidden_size, cfg.intermediate_size, bias=False)
        self.down_proj = nn.Linear(cfg.intermediate_size, cfg.hidden_size, bias=False)

    def forward(self, x):  # OpSlot("swiglu_mlp", "standard") call site (:123-124)
        return F.swiglu_mlp(self, x)


class Qwen3Attention(nn.Module):
    def __init__(self, cfg: Qwen3Config, layer_idx: int):
        super().__init__()
        self.cfg, self.layer_idx = cfg, layer_idx
        D = cfg.head_dim
        self.q_proj = nn.Linear(cfg.hidden_size, cfg.num_attention_heads * D, bias=False)
        self.k_proj = nn.Linear(cfg.hidden_size, cfg.num_key_value_heads * D, bias=False)
        self.v_proj = nn.Linear(cfg.hidden_size, cfg.num_key_value_heads * D, bias=False)
        self.o_proj = nn.Linear(cfg.num_attention_heads * D, cfg.hidden_size, bias=False)
        self.q_norm = Qwen3RMSNorm(D, cfg.rms_norm_eps)
        self.k_norm = Qwen3RMSNorm(D, cfg.rms_norm_eps)
        self.scaling = D**-0.5
        # gradient checkpointing: {forward id: (o, lse)} of first forwards, each consumed by its own recompute. The id
        # travels through the checkpointed call's arguments, so a recompute can only ever pick up the (o, lse) of the
        # forward it re-runs (two forwards before the first backward included).
        self.keep_attention = False
        self._stash: dict = {}

    def forward(self, x, cos, sin, cu_seqlens, max_seqlen, sp_group=None, fwd_id=None):
        # x: [T_local, hidden]; cos/sin: [T_local, D]
        cfg = self.cfg
        T, D = x.shape[0], cfg.head_dim
        q = self.q_proj(x).view(T, cfg.num_attention_heads, D)
        k = self.k_proj(x).view(T, cfg.num_key_value_heads, D)
        v = self.v_proj(x).view(T, cfg.num_key_value_heads, D)
        # q_norm / k_norm + apply_rotary_pos_emb (:305-310) in one pass
        out = None
        if sp_group is not None:
            from . import ulysses as U

            P = torch.distributed.get_world_size(sp_group)
            if P <= cfg.num_key_value_heads:  # write q, k straight into the Ulysses send staging (copy-free exchange)
                out = tuple(U.staging_views([tuple(q.shape), tuple(k.shape), tuple(v.shape)], q.dtype, sp_group)[:2])
        q, k = F.qknorm_rope(q, k, self.q_norm.weight, self.k_norm.weight, cos, sin, cfg.rms_norm_eps, out=out)
        if sp_group is not None:
            if P > cfg.num_key_value_heads:  # KV head replication (ops/kernels/attention/__init__.py:245-255)
                k = torch.repeat_interleave(k, P // cfg.num_key_value_heads, dim=1)
                v = torch.repeat_interleave(v, P // cfg.num_key_value_heads, dim=1)
            q, k, v = U.gather_seq_scatter_heads_qkv(q, k, v, seq_dim=0, head_dim=1, group=sp_group)
        keep = self.keep_attention and fwd_id is not None and torch.is_grad_enabled()
        st = self._stash.pop(fwd_id, None) if keep else None
        if st is not None and st[0].shape == (q.shape[0], q.shape[1], q.shape[2]) and st[1].shape[-1] == q.shape[0]:
            # recompute pass of this very forward: reuse the (deterministic) result instead of relaunching the kernel
            o = flash_attn_varlen(q, k, v, cu_seqlens, max_seqlen, self.scaling, True, replay=st)
        else:
            o, lse = flash_attn_varlen(q, k, v, cu_seqlens, max_seqlen, self.scaling, True, return_lse=True)
            if keep and self.training:
                self._stash[fwd_id] = (o.detach(), lse)
        if sp_group is not None:
            o = U.gather_heads_scatter_seq(o, head_dim=1, seq_dim=0, group=sp_group)
        return self.o_proj(o.reshape(T, -1))


class Qwen3DecoderLayer(nn.Module):
    def __init__(self, cfg: Qwen3Config, layer_idx: int):
        super().__init__()
        self.self_attn = Qwen3Attention(cfg, layer_idx)
        self.mlp = Qwen3MLP(cfg)
        self.input_layernorm = Qwen3RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.post_attention_layernorm = Qwen3RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.fuse_add_norm = os.environ.get("VB200_FUSE_ADD_NORM", "1") == "1" and cfg.hidden_size in F.FUSED_ADD_NORM_COLS

    def forward(self, h, cos, sin, cu_seqlens, max_seqlen, sp_group=None, fwd_id=None):
        a = self.self_attn(self.input_layernorm(h), cos, sin, cu_seqlens, max_seqlen, sp_group, fwd_id)
        if self.fuse_add_norm:  # residual add fused into the post-attention RMSNorm
            n = self.post_attention_layernorm
            x, h = F.fused_add_rms_norm(a, h, n.weight, n.variance_epsilon)
            return F.swiglu_mlp_residual(self.mlp, x, h)  # second residual add inside the down-projection GEMM
        h = h + a
        return h + self.mlp(self.post_attention_layernorm(h))


class Qwen3Model(nn.Module):
    def __init__(self, cfg: Qwen3Config):
        super().__init__()
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([Qwen3DecoderLayer(cfg, i) for i in range(cfg.num_hidden_layers)])
        self.norm = Qwen3RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)


class Qwen3ForCausalLM(nn.Module):
    _no_split_modules = ["Qwen3DecoderLayer"]

    def __init__(self, cfg: Qwen3Config):
        super().__init__()
        self.config = cfg
        self.model = Qwen3Model(cfg)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        if cfg.tie_word_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight
        self.gradient_checkpointing = False
        self.loss_impl = "fused"  # "fused": lm_head folded into the loss kernel's chunk loop; "logits": eager form
        self.keep_attention_in_checkpoint = True  # keep (o, lse) resident instead of recomputing attention
        self._fwd_counter = 0
        self.sp_group = None
        inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float() / cfg.head_dim))
        self.register_buffer("inv_freq", inv, persistent=False)

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None):
        self.gradient_checkpointing = True

    @torch.no_grad()
    def init_weights(self, seed: int = 0):
        """HF ``_init_weights`` semantics: N(0, initializer_range) for Linear/Embedding, ones for norms. Works on plain
        tensors and on FSDP2 DTensor shards (each rank fills its local shard; meta-init path of build_parallelize_model)."""
        dev = self.lm_head.weight.device
        g = torch.Generator(device=dev).manual_seed(seed + (_rank_seed() if _is_dtensor(self.lm_head.weight) else 0))
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                _local(m.weight).normal_(0.0, self.config.initializer_range, generator=g)
            elif isinstance(m, Qwen3RMSNorm):
                _local(m.weight).fill_(1.0)
        cfg = self.config
        self.inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64, device=dev).float() / cfg.head_dim))

    def rotary(self, position_ids: torch.Tensor, dtype: torch.dtype):
        """Qwen3RotaryEmbedding.forward (:181-192): fp32 angles, cos/sin cast to the activation dtype."""
        inv = self.inv_freq.to(device=position_ids.device, dtype=torch.float32)
        freqs = position_ids.reshape(-1).float()[:, None] * inv[None, :]
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.cos().to(dtype), emb.sin().to(dtype)

    def forward(self, input_ids, position_ids, cu_seqlens, max_seqlen, labels=None, shift_labels=None):
        """input_ids/position_ids/labels: [1, T] (packed row, local SP slice when Ulysses is on).

        ``cu_seqlens`` always describes the FULL packed row (data_collator.py:336-389 computes it
        before slicing), which is what attention sees after the Ulysses gather.
        """
        h = self.model.embed_tokens(input_ids.reshape(-1))
        cos, sin = self.rotary(position_ids, h.dtype)
        self._fwd_counter += 1
        ckpt = self.gradient_checkpointing and self.training
        fid = self._fwd_counter
        for layer in self.model.layers:
            att = layer.self_attn
            att.keep_attention = ckpt and self.keep_attention_in_checkpoint
            for old in [k for k in att._stash if k < fid - 4]:  # forwards that never got a backward (bounded memory)
                del att._stash[old]
            if ckpt:
                h = checkpoint(layer, h, cos, sin, cu_seqlens, max_seqlen, self.sp_group, fid, use_reentrant=False)
            else:
                h = layer(h, cos, sin, cu_seqlens, max_seqlen, self.sp_group)
        h = self.model.norm(h)
        if labels is None and shift_labels is None:
            return self.lm_head(h)
        # ForCausalLMLoss (cross_entropy/__init__.py:180-221): shift unless already shifted (SP), then the bound
        # cross_entropy_fn. "fused" hands hidden states + lm_head weight to the kernel (the reference's liger path);
        # "logits" materialises bf16 logits and upcasts them like the reference's eager path.
        if shift_labels is None:
            shift_labels = Fnn.pad(labels, (0, 1), value=-100)[..., 1:].contiguous()
        if self.loss_impl == "fused":
            loss, _ = b200_cross_entropy(None, shift_labels, self.config.vocab_size, hidden_states=h,
                                         weights=self.lm_head.weight)
        else:
            loss, _ = b200_cross_entropy(self.lm_head(h).float().view(-1, self.config.vocab_size), shift_labels,
                                         self.config.vocab_size)
        return loss


def _is_dtensor(t) -> bool:
    from torch.distributed._tensor import DTensor

    return isinstance(t, DTensor)


def _local(t):
    return t.to_local() if _is_dtensor(t) else t


def _rank_seed() -> int:
    import torch.distributed as dist

    return 7919 * (dist.get_rank() + 1) if dist.is_initialized() else 0


def flops_per_token(cfg: Qwen3Config, seq_lens: list[int]) -> float:
    """Reference MFU accounting: 6*N_dense*tokens + 12*sum(L^2)*head_dim*heads*layers
    (veomni/utils/count_flops.py:221-253; N_dense counts lm_head but not the embedding)."""
    n_dense = cfg.num_params() - cfg.vocab_size * cfg.hidden_size * (0 if cfg.tie_word_embeddings else 1)
    tokens = sum(seq_lens)
    attn = 12 * sum(s * s for s in seq_lens) * cfg.head_dim * cfg.num_attention_heads * cfg.num_hidden_layers
    return (6 * n_dense * tokens + attn) / tokens
