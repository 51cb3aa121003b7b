"""One tiny forward+backward of the host Qwen3 caller on cuda:0 (used by __graft_entry__.smoke)."""
import torch


def run(dev: torch.device) -> float:
    from .host_qwen3 import Qwen3Config, Qwen3ForCausalLM

    cfg = Qwen3Config(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=64)
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
    val = float(loss)
    assert 5.0 < val < 8.0, f"unexpected smoke loss {val} (ln(512) = 6.24 expected for random init)"
    return val
