"""FSDP2 wrapping of a model the way VeOmni does it, with the NVLink collectives installed.

Mirror of ``build_parallelize_model`` / ``parallelize_model_fsdp2``
(veomni/distributed/torch_parallelize.py:419-480, :76-416) for the dense data-parallel case:
bottom-up ``fully_shard`` of every ``_no_split_modules`` class with
``MixedPrecisionPolicy(param_dtype=bf16, reduce_dtype=fp32)`` (:198-205, defaults
veomni/arguments/arguments_types.py:248-255) and ``reshard_after_forward`` as given (:286), then the
root without an explicit ``reshard_after_forward`` (:334-344), fp32 master weights (:442-443) and
per-layer gradient checkpointing (:445-456).  With the reference installed the same effect is obtained
by calling ``veomni_b200.fsdp_comm.install_fsdp_comm(model)`` after its own
``build_parallelize_model``.
"""

from __future__ import annotations

import torch
import torch.distributed as dist
from torch.distributed.device_mesh import DeviceMesh, init_device_mesh
from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

from .fsdp_comm import install_fsdp_comm


def build_parallelize_model(
    model: torch.nn.Module,
    enable_reshard_after_forward: bool = True,
    param_dtype: torch.dtype = torch.bfloat16,
    reduce_dtype: torch.dtype = torch.float32,
    enable_gradient_checkpointing: bool = True,
    basic_modules: list[str] | None = None,
    mesh: DeviceMesh | None = None,
    b200_comm: bool = True,
    comm_ctas: int = 32,
) -> torch.nn.Module:
    if mesh is None:
        mesh = init_device_mesh("cuda", (dist.get_world_size(),), mesh_dim_names=("dp_shard",))
    model = model.float()  # fp32 master weights; FSDP2 casts to param_dtype for compute
    if enable_gradient_checkpointing and hasattr(model, "gradient_checkpointing_enable"):
        model.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    targets = set(getattr(model, "_no_split_modules", None) or []) | set(basic_modules or [])
    mp = MixedPrecisionPolicy(param_dtype=param_dtype, reduce_dtype=reduce_dtype)
    mods = [(fqn, m) for fqn, m in model.named_modules() if m.__class__.__name__ in targets]
    for fqn, m in sorted(mods, key=lambda t: -t[0].count(".")):  # submodules first
        fully_shard(m, mesh=mesh, mp_policy=mp, reshard_after_forward=enable_reshard_after_forward)
    fully_shard(model, mesh=mesh, mp_policy=mp)
    if b200_comm:
        model._vb200_symm = install_fsdp_comm(model, mesh.get_group(), num_ctas=comm_ctas)
    # the reference binds the FSDP2-aware clip onto the model (torch_parallelize.py:412-414)
    import types

    from .clip_grad_norm import clip_grad_norm

    model.clip_grad_norm_ = types.MethodType(clip_grad_norm, model)
    return model
