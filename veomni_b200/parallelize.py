"""FSDP2 (+EP, +HSDP) wrapping of a model the way VeOmni does it, with the NVLink collectives installed.

Mirror of ``build_parallelize_model`` (veomni/distributed/torch_parallelize.py:419-480) and
``parallelize_model_fsdp2`` (:76-416):

* world size 1 (``fsdp_enabled`` false, :438-440,465): no FSDP wrapping at all — fp32 weights, per-layer gradient
  checkpointing, the clip method bound; the caller pairs it with ``B200AdamW(master_weights=True)``;
* expert parallelism (:129-160): ``ParallelPlan.apply`` slices the expert tensors ``[E, ...] -> [E/EP, ...]``, the
  expert modules are ``fully_shard``-ed on their own ``ep_fsdp`` mesh along dim 1 (:239-259,297-301) with the gradient
  divide factor of the whole world (:306-313), and get their own NVLink comm on that group;
* every ``_no_split_modules`` / ``basic_modules`` class bottom-up with
  ``MixedPrecisionPolicy(param_dtype=bf16, reduce_dtype=fp32)`` (:198-205; defaults arguments_types.py:248-255) and
  ``reshard_after_forward`` as given (:197), the root without an explicit ``reshard_after_forward`` (:334-344);
* HSDP: a 2-D ``(dp_replicate, dp_shard_sp)`` mesh when ``ParallelState.dp_replicate_size > 1``
  (parallel_state.py:253-261) — FSDP2 then all-reduces the reduce-scatter output over the replica group with NCCL,
  which is the one place ``north_star`` keeps NCCL on the data path;
* manual forward / backward prefetch lists when EP is on (:346-365);
* ``init_device="meta"`` (:367-374): materialise on the GPU after sharding and call ``init_weights``.

With the reference installed the same effect is obtained by calling
``veomni_b200.fsdp_comm.install_fsdp_comm(model)`` after its own ``build_parallelize_model``.
"""

from __future__ import annotations

import types

import torch
import torch.distributed as dist
from torch.distributed._tensor import Shard
from torch.distributed.device_mesh import DeviceMesh, init_device_mesh
from torch.distributed.fsdp import FSDPModule, MixedPrecisionPolicy, fully_shard

from .fsdp_comm import install_fsdp_comm
from .parallel_state import get_parallel_state


def _bind_clip(model: torch.nn.Module) -> None:
    # the reference binds the FSDP2-aware clip onto the model (torch_parallelize.py:412-414)
    from .clip_grad_norm import clip_grad_norm

    model.clip_grad_norm_ = types.MethodType(clip_grad_norm, model)


def _sorted_submodules_first(mods: list[tuple[str, torch.nn.Module]]):
    return sorted(mods, key=lambda t: -t[0].count("."))


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
    rs_mode: str | None = None,
    fuse_copy_out: bool = True,
    init_device: str | None = None,
    enable_forward_prefetch: bool = True,
    misc_bytes: int = 256 << 20,
) -> torch.nn.Module:
    ps = get_parallel_state()
    world = dist.get_world_size() if dist.is_initialized() else 1
    model = model.float()  # fp32 master weights; FSDP2 (or the master-weight optimizer) casts to param_dtype for compute
    if enable_gradient_checkpointing and hasattr(model, "gradient_checkpointing_enable"):
        model.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
    if world == 1 and mesh is None:
        if init_device == "meta":
            raise ValueError("Only FSDP training supports `init_device=meta`.")  # torch_parallelize.py:438-440
        _bind_clip(model)
        return model

    # ---- meshes -------------------------------------------------------------------------------------------------
    if mesh is None:
        if ps.device_mesh is not None:
            mesh = ps.fsdp_mesh  # 2-D (dp_replicate, dp_shard[_sp]) under HSDP
        else:
            mesh = init_device_mesh("cuda" if torch.cuda.is_available() else "cpu", (world,), mesh_dim_names=("dp_shard",))
    shard_group = mesh.get_group(mesh.ndim - 1)  # the reduce-scatter / all-gather group (last mesh dim)
    mp = MixedPrecisionPolicy(param_dtype=param_dtype, reduce_dtype=reduce_dtype)
    fsdp_kwargs = dict(mesh=mesh, mp_policy=mp, reshard_after_forward=enable_reshard_after_forward)

    # ---- step 1: expert parallelism (slice, then shard the experts over ep_fsdp along dim 1) ---------------------
    ep_modules: dict[str, torch.nn.Module] = {}
    ep_kwargs = None
    if ps.ep_enabled:
        plan = model.get_parallel_plan()
        if plan is None:
            raise AssertionError("ExtraParallel needs parallel plan defined in the model!")
        infos = plan.apply(model, ps.ep_size, ps.ep_rank)
        model._fqn2spec_info = infos
        sharded = {fqn.rsplit(".", 1)[0] for fqn, info in infos.items() if isinstance(info.placement, Shard)}
        ep_modules = {fqn: m for fqn, m in model.named_modules() if fqn in sharded}
        ep_kwargs = dict(fsdp_kwargs)
        ep_kwargs["mesh"] = ps.ep_fsdp_device_mesh["ep_fsdp"]
        ep_kwargs["shard_placement_fn"] = lambda param: Shard(1)

    # ---- step 2: bottom-up fully_shard ---------------------------------------------------------------------------
    targets = set(getattr(model, "_no_split_modules", None) or []) | set(basic_modules or [])
    mods = [(fqn, m) for fqn, m in model.named_modules() if m.__class__.__name__ in targets]
    blocks = []
    for fqn, m in _sorted_submodules_first(mods):
        m._fsdp_modules = []
        for efqn, em in ep_modules.items():
            if efqn.startswith(fqn) and not isinstance(em, FSDPModule):
                fully_shard(em, **ep_kwargs)
                em.set_gradient_divide_factor(float(ps.extra_parallel_gradient_divide_factor("ep")))
                m._fsdp_modules.append(em)
        if not isinstance(m, FSDPModule):
            fully_shard(m, **fsdp_kwargs)
            m._fsdp_modules.append(m)
        blocks.append((fqn, m))
    root_kwargs = {k: v for k, v in fsdp_kwargs.items() if k != "reshard_after_forward"}
    fully_shard(model, **root_kwargs)

    # ---- explicit forward prefetch, dense path ------------------------------------------------------------------------
    # FSDP2's implicit forward prefetch is nothing but the host running ahead: unit i+1's all-gather is enqueued by
    # pre_forward(i+1), i.e. after the host has finished enqueuing all of unit i's kernels. With ~1.8 ms of GPU work per
    # Qwen3-8B layer forward and FSDP2's hook overhead on top of the launches, the host is barely ahead, and every all-gather
    # that is enqueued late is exposed in full (the N>=2 step lost 30-45 ms that way). Asking unit i to issue unit i+1's
    # all-gather in its own pre_forward (before its kernels are enqueued) makes the overlap independent of host timing; the
    # cost is one more unsharded unit resident (0.4 GB). Numerics are untouched.
    if not ps.ep_enabled and enable_forward_prefetch and len(blocks) > 1:
        ordered = [m for _f, m in sorted(blocks, key=lambda t: [int(x) if x.isdigit() else x for x in t[0].split(".")])]
        for cur, nxt in zip(ordered, ordered[1:]):
            cur.set_modules_to_forward_prefetch([nxt])
        model.set_modules_to_forward_prefetch([ordered[0]])  # root (embedding / lm_head / final norm) -> first block

    # ---- manual prefetch (EP, as the reference) ------------------------------------------------------------------
    if ps.ep_enabled and enable_forward_prefetch:
        ordered = [m for _f, m in sorted(blocks, key=lambda t: [int(x) if x.isdigit() else x for x in t[0].split(".")])]
        for cur, nxt in zip(ordered, ordered[1:]):
            cur.set_modules_to_forward_prefetch(list(reversed(nxt._fsdp_modules)))
        rev = list(reversed(ordered))
        for cur, prv in zip(rev, rev[1:]):
            cur.set_modules_to_backward_prefetch(list(reversed(prv._fsdp_modules)))

    # ---- meta init --------------------------------------------------------------------------------------------------
    if init_device == "meta":
        on_gpu = mesh.device_type == "cuda"  # CPU meshes (gloo) only in the host-logic tests
        model.to_empty(device=torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu"))
        model.init_weights()

    # ---- the NVLink collectives ---------------------------------------------------------------------------------------
    if b200_comm:
        dense = [m for m in model.modules() if isinstance(m, FSDPModule) and not any(m is e for e in ep_modules.values())]
        experts = [m for m in ep_modules.values() if isinstance(m, FSDPModule)]
        if dist.get_world_size(shard_group) > 1:
            model._vb200_symm = install_fsdp_comm(model, shard_group, num_ctas=comm_ctas, rs_mode=rs_mode,
                                                  fuse_copy_out=fuse_copy_out, modules=dense, misc_bytes=misc_bytes)
        if experts and ps.ep_fsdp_size > 1:
            model._vb200_symm_ep = install_fsdp_comm(model, ps.ep_fsdp_device_mesh["ep_fsdp"].get_group(), num_ctas=comm_ctas,
                                                     rs_mode=rs_mode, fuse_copy_out=fuse_copy_out, modules=experts)
    if ps.ep_enabled:
        # fully_shard replaced every parameter by a new DTensor nn.Parameter: put the ExtraParallel tags back, the
        # EP-aware gradient clip tells expert parameters apart by them (fsdp2/clip_grad_norm.py:60-75)
        for fqn, p in model.named_parameters():
            if fqn in model._fqn2spec_info:
                p.spec_info = model._fqn2spec_info[fqn]
    _bind_clip(model)
    return model
