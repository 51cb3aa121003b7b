"""Generate tests/golden/wrap_structure.json: what the REFERENCE's ``build_parallelize_model`` does to a toy Qwen3-MoE.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_wrap_structure.py

For each (world, ep_size) in WORLDS the reference, unmodified, runs on CPU / gloo (shim of SURVEY.md Appendix B: gloo
backend, ``torch.cpu`` memory-stat stubs): ``init_parallel_state`` -> ``build_foundation_model`` (toy ``qwen3_moe``, meta
device, all ops eager) -> ``build_parallelize_model(init_device="meta", ...)`` exactly as ``BaseTrainer`` calls it
(veomni/trainer/base.py:387-404). ``describe`` (shared with tests/test_parallel_host.py, which applies it to
``veomni_b200.parallelize.build_parallelize_model`` on the host model of the same shape) then records, per rank:

* which modules became FSDP2 units, in wrap order, and each unit's mesh (dim names + sizes), reshard-after-forward flag,
  gradient divide factor, mixed-precision dtypes and parameter list with placements and local shapes;
* every unit's explicit forward / backward prefetch targets (torch_parallelize.py:346-365);
* which parameter names ``model._fqn2spec_info`` marks as ExtraParallel-sliced (the EP-aware clip keys on it).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = "/root/reference"
# (world, ep_size, dp_replicate): ep_fsdp = 1, ep_fsdp = 2, dense-only, HSDP (2 replicas x 2 shards)
WORLDS = [(2, 2, 1), (4, 2, 1), (2, 1, 1), (4, 1, 2)]
TOY = dict(architectures=["Qwen3MoeForCausalLM"], model_type="qwen3_moe", hidden_size=64, intermediate_size=128,
           moe_intermediate_size=32, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
           num_experts=8, num_experts_per_tok=2, norm_topk_prob=True, decoder_sparse_step=1, mlp_only_layers=[],
           vocab_size=256, max_position_embeddings=512, rms_norm_eps=1e-6, rope_theta=1000000.0, tie_word_embeddings=False,
           hidden_act="silu", initializer_range=0.02, attention_bias=False, attention_dropout=0.0, output_router_logits=False,
           router_aux_loss_coef=0.001, use_sliding_window=False, sliding_window=None, dtype="bfloat16")


def describe(model) -> dict:
    """Structure of an FSDP2-wrapped model, in plain JSON types (works on the reference's model and on ours)."""
    import torch
    from torch.distributed._tensor import DTensor
    from torch.distributed.fsdp import FSDPModule

    names = {id(m): n for n, m in model.named_modules()}
    pnames = {id(p): n for n, p in model.named_parameters()}
    units = []
    for n, m in model.named_modules():
        if not isinstance(m, FSDPModule):
            continue
        st = m._get_fsdp_state()
        pg = st._fsdp_param_group
        u = {"module": n, "cls": next(c.__name__ for c in type(m).__mro__ if not c.__name__.startswith("FSDP")),
             "auto_reshard_after_forward": bool(getattr(st, "_auto_reshard_after_forward", False))}
        if pg is not None:
            mesh = pg.mesh_info.mesh
            u["mesh"] = [[d, int(mesh.size(i))] for i, d in enumerate(mesh.mesh_dim_names or ("?",) * mesh.ndim)]
            post = pg.post_forward_mesh_info
            u["reshard_after_forward"] = post is not None
            factor = getattr(pg, "gradient_divide_factor", None)
            u["gradient_divide_factor"] = None if factor is None else float(factor)
            u["param_dtype"] = str(pg.mp_policy.param_dtype)
            u["reduce_dtype"] = str(pg.mp_policy.reduce_dtype)
            u["params"] = [[pnames.get(id(p.sharded_param), "?"), [str(pl) for pl in p.sharded_param.placements],
                            list(p.sharded_param.to_local().shape)] for p in pg.fsdp_params]
        u["forward_prefetch"] = [names.get(id(s._modules[0] if hasattr(s, "_modules") else None), "?")
                                 for s in getattr(st, "_states_to_forward_prefetch", [])]
        u["backward_prefetch"] = [names.get(id(s._modules[0] if hasattr(s, "_modules") else None), "?")
                                  for s in getattr(st, "_states_to_backward_prefetch", [])]
        units.append(u)
    infos = getattr(model, "_fqn2spec_info", None) or {}
    tagged = sorted(k for k, v in infos.items() if type(getattr(v, "placement", None)).__name__ == "Shard")
    dt = sorted({str(p.dtype) for p in model.parameters()})
    return {"units": units, "ep_sliced_fqns": tagged, "param_dtypes": dt,
            "all_dtensor": all(isinstance(p, DTensor) for p in model.parameters()),
            "has_clip": hasattr(model, "clip_grad_norm_"), "grad_ckpt": bool(getattr(model, "is_gradient_checkpointing", False)
                                                                               or getattr(model, "gradient_checkpointing", False))}


def _worker(rank: int, world: int, ep: int, rep: int, store: str, cfg_dir: str, out: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      OMP_NUM_THREADS="1")
    sys.path.insert(0, REF)
    sys.path.insert(0, f"{REF}/tests")
    import torch
    import torch.distributed as dist

    torch.set_num_threads(1)
    import veomni.utils.device as dev

    dev.get_dist_comm_backend = lambda: "gloo"
    for n, v in dict(get_device_name=lambda *a: "CPU", max_memory_allocated=lambda *a: 0, max_memory_reserved=lambda *a: 0,
                     memory_stats=lambda *a: {"num_alloc_retries": 0}, memory_allocated=lambda *a: 0,
                     memory_reserved=lambda *a: 0, empty_cache=lambda *a: None,
                     reset_peak_memory_stats=lambda *a: None).items():
        setattr(torch.cpu, n, v)
    dist.init_process_group("gloo", store=dist.FileStore(store, world), rank=rank, world_size=world)
    from tools.training_utils import make_eager_ops_config  # reference test helper
    from veomni.distributed.parallel_state import init_parallel_state
    from veomni.distributed.torch_parallelize import build_parallelize_model
    from veomni.models import build_foundation_model

    init_parallel_state(dp_size=world, dp_replicate_size=rep, dp_shard_size=world // rep, ulysses_size=1, dp_mode="fsdp2",
                        device_type="cpu", extra_parallel_sizes=(ep,))
    model = build_foundation_model(config_path=cfg_dir, torch_dtype="float32", attn_implementation="sdpa", init_device="meta",
                                   ops_implementation=make_eager_ops_config())
    cpu_load = getattr(model.get_parallel_plan(), "cpu_load_param_name", None) if hasattr(model, "get_parallel_plan") else None
    model = build_parallelize_model(
        model, init_device="meta", weights_path=None, enable_reshard_after_forward=True,
        enable_gradient_checkpointing=True, basic_modules=list(set(getattr(model, "_no_split_modules", None) or [])),
        enable_reentrant=False, enable_forward_prefetch=True, enable_fsdp_offload=False,
        broadcast_model_weights_from_rank0=False, cpu_load_param_name=cpu_load)
    res = describe(model)
    with open(f"{out}/r{rank}.json", "w") as fh:
        json.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp

    result = {"toy_config": TOY, "runs": []}
    with tempfile.TemporaryDirectory() as d:
        cfg = Path(d) / "cfg"
        cfg.mkdir()
        (cfg / "config.json").write_text(json.dumps(TOY))
        for world, ep, rep in WORLDS:
            out = Path(d) / f"w{world}e{ep}r{rep}"
            out.mkdir()
            mp.spawn(_worker, args=(world, ep, rep, str(out / "store"), str(cfg), str(out)), nprocs=world, join=True)
            ranks = [json.loads((out / f"r{r}.json").read_text()) for r in range(world)]
            result["runs"].append({"world": world, "ep_size": ep, "dp_replicate": rep, "ranks": ranks})
            print(f"world {world} ep {ep} replicas {rep}: {len(ranks[0]['units'])} FSDP units, ep-sliced {len(ranks[0]['ep_sliced_fqns'])}")
    (HERE / "wrap_structure.json").write_text(json.dumps(result, indent=1))
    print("wrote", HERE / "wrap_structure.json")


if __name__ == "__main__":
    main()
