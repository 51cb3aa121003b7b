"""One tiny forward+backward of the host Qwen3 caller on cuda:0 (used by __graft_entry__.smoke)."""
import torch


def run(dev: torch.device) -> float:
    from .host_qwen3 import Qwen3Config, Qwen3ForCausalLM

    # head_dim 128 (Qwen3-8B's) so that the tcgen05 attention kernels are the ones that run
    cfg = Qwen3Config(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=128)
    model = Qwen3ForCausalLM(cfg).to(dev).to(torch.bfloat16)
    model.init_weights(seed=0)
    model.gradient_checkpointing_enable()
    model.train()
    lens = [70, 58]
    T = sum(lens)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 512, (1, T), generator=g).to(dev)
    pos = torch.cat([torch.arange(n) for n in lens])[None].to(dev)
    cu = torch.tensor([0, 70, 128], dtype=torch.int32, device=dev)
    loss = model(ids, pos, cu, max(lens), labels=ids)
    loss.backward()
    val = float(loss.detach())
    assert 5.0 < val < 8.0, f"unexpected smoke loss {val} (ln(512) = 6.24 expected for random init)"
    return val


def run_moe(dev: torch.device) -> None:
    """One fused-MoE forward + backward (routing kernels + tcgen05 GroupGEMM fc1 / fc2 / dgrad / wgrad)."""
    from .moe import fused_moe_forward

    T, E, K, H, I = 256, 8, 2, 256, 128
    g = torch.Generator().manual_seed(1)
    hs = (0.1 * torch.randn(T, H, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
    w1 = (0.1 * torch.randn(E, 2 * I, H, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
    w2 = (0.1 * torch.randn(E, H, I, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
    rw, idx = torch.topk(torch.softmax(torch.randn(T, E, generator=g), -1), K, dim=-1)
    out = fused_moe_forward(E, rw.to(torch.bfloat16).to(dev), idx.to(dev), hs, None, None, w2, fc1_1_2_weight=w1)
    out.float().square().mean().backward()
    assert torch.isfinite(out.float()).all() and torch.isfinite(w1.grad.float()).all()
