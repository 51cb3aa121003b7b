"""Oracle for the memory-bound block ops (RMSNorm, RoPE, q/k-norm+RoPE, SwiGLU).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  All functions take and return torch CPU
tensors; ``dtype`` semantics follow the reference exactly (the reference computes these ops in
the activation dtype with the explicit fp32 up-casts shown below).
"""

from __future__ import annotations

import torch


def rms_norm(hidden_states: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """Qwen3RMSNorm.forward eager body.

    Reference: veomni/models/transformers/qwen3/generated/patched_modeling_qwen3_gpu.py:93-97
    (identical text in qwen3_moe/generated/patched_modeling_qwen3_moe_gpu.py, Qwen3MoeRMSNorm).
    """
    input_dtype = hidden_states.dtype
    h = hidden_states.to(torch.float32)
    variance = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(variance + eps)
    return weight * h.to(input_dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """Reference: patched_modeling_qwen3_gpu.py:196-200."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2 :]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, unsqueeze_dim: int = 1):
    """Reference: patched_modeling_qwen3_gpu.py:208-223 (eager branch)."""
    cos = cos.unsqueeze(unsqueeze_dim)
    sin = sin.unsqueeze(unsqueeze_dim)
    q_embed = (q * cos) + (rotate_half(q) * sin)
    k_embed = (k * cos) + (rotate_half(k) * sin)
    return q_embed, k_embed


def rotary_cos_sin(position_ids: torch.Tensor, head_dim: int, rope_theta: float, dtype: torch.dtype):
    """Qwen3RotaryEmbedding.forward for the default rope type (attention_scaling = 1).

    Reference: patched_modeling_qwen3_gpu.py:160-192 (inv_freq = 1/theta^(2i/d) :172-176;
    freqs = inv_freq x position, emb = cat(freqs, freqs), cos/sin in fp32 then cast :181-192).
    position_ids: [B, S] integer tensor. Returns cos, sin of shape [B, S, head_dim].
    """
    inv_freq = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float) / head_dim))
    inv_freq_expanded = inv_freq[None, :, None].float().expand(position_ids.shape[0], -1, 1)
    position_ids_expanded = position_ids[:, None, :].float()
    freqs = (inv_freq_expanded.float() @ position_ids_expanded.float()).transpose(1, 2)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def qknorm_rope(q, k, wq, wk, cos, sin, eps: float):
    """q_norm / k_norm over head_dim then RoPE, as Qwen3Attention.forward does.

    Reference: patched_modeling_qwen3_gpu.py:305-310.  q: [T, Hq, D], k: [T, Hk, D],
    cos/sin: [T, D].  Returns (q, k) in the same [T, H, D] layout.
    """
    qn = rms_norm(q, wq, eps)
    kn = rms_norm(k, wk, eps)
    # reference layout is [B, H, S, D] with cos/sin [B, S, D] unsqueezed at dim 1
    qe, ke = apply_rotary_pos_emb(qn.transpose(0, 1)[None], kn.transpose(0, 1)[None], cos[None], sin[None])
    return qe[0].transpose(0, 1).contiguous(), ke[0].transpose(0, 1).contiguous()


def silu_mul(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """``act_fn(gate) * up`` of Qwen3MLP.forward (patched_modeling_qwen3_gpu.py:126) and of the MoE
    expert body (veomni/ops/kernels/moe/group_gemm.py:300-304, distributed/moe/moe_layer.py:343-346)."""
    return torch.nn.functional.silu(gate) * up


def swiglu_mlp(x, gate_w, up_w, down_w):
    """Qwen3MLP.forward (patched_modeling_qwen3_gpu.py:121-127)."""
    F = torch.nn.functional
    return F.linear(silu_mul(F.linear(x, gate_w), F.linear(x, up_w)), down_w)
