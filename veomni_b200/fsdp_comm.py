"""FSDP2 parameter all-gather / gradient reduce-scatter on the NVLink pull kernels.

Drop-in boundary (SURVEY.md §8(b) "FSDP comm"): PyTorch's official hook
``FSDPModule.set_custom_all_gather(AllGather)`` / ``set_custom_reduce_scatter(ReduceScatter)``
(torch/distributed/fsdp/_fully_shard/_fully_shard.py:458-482; ABCs in _fsdp_api.py:56-127), which
VeOmni reaches through ``build_parallelize_model`` -> ``fully_shard``
(veomni/distributed/torch_parallelize.py:289-344).

* ``allocate`` hands FSDP2 a tensor that aliases this rank's symmetric region, so FSDP2's own copy-in
  (``fsdp::all_gather_copy_in`` / ``chunk_cat``) writes the local shard straight into peer-visible
  memory;
* ``__call__`` runs one kernel on FSDP2's all-gather / reduce-scatter stream (the current stream at
  call time): peers' shards are pulled over NVLink with 16-byte loads; the reduce-scatter sums in
  fp32 in rank order 0..N-1 (deterministic) and applies the AVG / pre-multiplied-SUM factor
  (``_get_gradient_divide_factors``, _fsdp_collectives.py:701-759) in the same pass.
"""

from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist
from torch.distributed.fsdp._fully_shard._fsdp_api import AllGather, ReduceScatter

from . import prof
from ._lib import VB200Error
from .symm import SymmetricMemory, get_symmetric_memory

CH_ALL_GATHER = 0
CH_REDUCE_SCATTER = 1


def _reduce_scale(op, world: int) -> float:
    """Scale applied to the fp32 sum for the ReduceOp FSDP2 passes."""
    ReduceOp = dist.ReduceOp
    if op == ReduceOp.SUM:
        return 1.0
    if op == ReduceOp.AVG:
        return 1.0 / world
    # PREMUL_SUM carries its factor in the op's pickled state: (RedOpType.PREMUL_SUM, factor)
    if op == ReduceOp.PREMUL_SUM:
        factor = op.__getstate__()[1]
        if isinstance(factor, torch.Tensor):
            factor = factor.item()
        return float(factor)
    raise VB200Error(f"unsupported reduce op for the B200 reduce-scatter: {op!r}")


# all-gather outputs (address) whose per-parameter pieces B200AllGather already delivered straight into the unsharded
# parameter tensors -> the tensors FSDP2's copy-out must *not* copy into again (consumed by _split_with_sizes_copy)
_SCATTERED: set[int] = set()
# the FSDPParams of the foreach_all_gather call in flight (set by _param_all_gather_inputs, consumed by __call__)
_AG_CALL: list = []


class B200AllGather(AllGather):
    """FSDP2 unit all-gather on the NVLink pull kernel.

    With ``fuse_copy_out`` (default) the kernel stores every peer's shard straight into the per-parameter unsharded
    tensors (``FSDPParam.all_gather_outputs``), which are allocated here — on the all-gather stream, exactly as FSDP2's
    own copy-out would size them — so ``foreach_all_gather_copy_out``'s ``split_with_sizes_copy`` pass
    (_fsdp_collectives.py:196-212,346-412; 14 ms of a 279 ms Qwen3-8B step at N=4) has nothing left to do.
    """

    def __init__(self, symm: SymmetricMemory, num_ctas: int = 32, fuse_copy_out: bool = True):
        self.symm = symm
        self.num_ctas = num_ctas
        self.fuse_copy_out = fuse_copy_out

    def allocate(self, size: Sequence[int], *, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        return self.symm.empty(tuple(int(s) for s in size), dtype, arena="fsdp_ag")

    def _scatter_table(self, output_tensor, input_tensor, world):
        """Parameter table for the fused copy-out, or None when this call must take the two-step path."""
        from torch.distributed.fsdp._fully_shard._fsdp_param import ShardedState

        params = list(_AG_CALL)
        _AG_CALL.clear()
        if not self.fuse_copy_out or not params:
            return None
        es = input_tensor.element_size()
        numels = []
        for p in params:
            if p.fsdp_placement.dim != 0 or len(p.all_gather_outputs) > 1:
                return None
            t = p._sharded_param_data if p.sharded_state == ShardedState.SHARDED else p._sharded_post_forward_param_data
            numels.append(t.numel())
        if sum(numels) != input_tensor.numel() or len(params) > 64:
            return None
        table, off = [], 0
        for p, n in zip(params, numels):
            # what foreach_all_gather_copy_out does first (:376-383); both calls are no-ops when it repeats them
            p.init_all_gather_outputs([n], [input_tensor.dtype], world, input_tensor.device)
            p.alloc_all_gather_outputs()
            dst = p.all_gather_outputs[0]
            if dst.dtype != input_tensor.dtype or dst.numel() != n * world:
                return None
            table += [off * es, n * es, dst.data_ptr()]
            off += n
        return table

    def __call__(self, output_tensor, input_tensor, group, async_op: bool = False):
        world = dist.get_world_size(group)
        if world != self.symm.world:
            raise VB200Error("all-gather group does not match the symmetric-memory group")
        if input_tensor.data_ptr() != output_tensor.data_ptr() + self.symm.rank * input_tensor.numel() * input_tensor.element_size():
            # FSDP2's copy-in always produces input = output[rank*n:(rank+1)*n]; keep the general case correct
            output_tensor.narrow(0, self.symm.rank * input_tensor.numel(), input_tensor.numel()).copy_(input_tensor)
        table = self._scatter_table(output_tensor, input_tensor, world)
        if table is None:
            with prof.span("fsdp_all_gather", input_tensor.numel() * input_tensor.element_size() * (world - 1)):
                self.symm.all_gather_inplace(output_tensor, input_tensor.numel(), CH_ALL_GATHER, self.num_ctas)
            return None
        with prof.span("fsdp_all_gather", input_tensor.numel() * input_tensor.element_size() * (world - 1)):
            self.symm.all_gather_scatter(output_tensor, input_tensor.numel(), table, CH_ALL_GATHER, self.num_ctas)
        _SCATTERED.add(output_tensor.data_ptr())
        return None


def _split_with_sizes_copy(all_gather_output, all_gather_input_split_sizes, dim=0, *, out):
    """CUDA implementation of ``fsdp::split_with_sizes_copy`` (_fsdp_collectives.py:196-212). When the all-gather kernel
    already stored the shards into ``out`` there is no copy left; the tensors were allocated on the all-gather stream and
    are read on this (the compute) stream, so the caching allocator is told about the second stream."""
    key = all_gather_output.data_ptr()
    if key in _SCATTERED:
        _SCATTERED.discard(key)
        cur = torch.cuda.current_stream()
        for o in out:
            o.record_stream(cur)
        return
    torch.split_with_sizes_copy(all_gather_output, all_gather_input_split_sizes, dim=dim, out=out)


# reduce-scatter inputs handed out by B200ReduceScatter.allocate, by address: lets the patched copy-in recognise them
_RS_INPUTS: dict[int, "B200ReduceScatter"] = {}


class B200ReduceScatter(ReduceScatter):
    """FSDP2 unit reduce-scatter on the NVLink kernels.

    ``mode="push"`` (default): the copy-in that precedes the collective is ours and does not copy — :func:`_copy_in` only
    remembers the unit's bf16 gradients; ``__call__`` then runs ONE kernel on the reduce-scatter stream that reads them in
    place, pushes every peer its chunk over NVLink and reduces the staged chunks in fp32 in rank order
    (``vb200_reduce_scatter_push_bf16``). ``mode="pull"``: a bf16 pack kernel on the compute stream + the pull-reduce
    kernel (the round-1 path). ``mode="f32"``: torch's fp32 ``chunk_cat`` copy-in + the fp32 pull-reduce. All three give
    bit-identical sums (fixed rank order; bf16 -> fp32 is exact).
    """

    def __init__(self, symm: SymmetricMemory, num_ctas: int = 32, pack_bf16: bool = True, mode: str | None = None):
        self.symm = symm
        self.num_ctas = num_ctas
        self.mode = mode if mode is not None else ("push" if pack_bf16 else "f32")
        if self.mode not in ("push", "pull", "f32"):
            raise VB200Error(f"unknown reduce-scatter mode {self.mode!r}")
        self.pack_bf16 = self.mode != "f32"
        self._next_is_input = True
        self._packed: dict[int, int] = {}  # input address -> bf16 elements per rank chunk (pull mode)
        self._pending: dict[int, tuple] = {}  # input address -> (gradients, plan, row) (push mode)

    def allocate(self, size: Sequence[int], *, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        # foreach_reduce (_fsdp_collectives.py:522-541) asks for the (N*chunk) input first and for the
        # (chunk) output second. Only the input must be peer-visible; the output backs the sharded
        # .grad for the rest of the step, so it comes from the ordinary caching allocator.
        shape = tuple(int(s) for s in size)
        if self._next_is_input:
            self._next_is_input = False
            t = self.symm.empty(shape, dtype, arena="fsdp_rs")
            _RS_INPUTS[t.data_ptr()] = self
            self._packed.pop(t.data_ptr(), None)
            self._pending.pop(t.data_ptr(), None)
            return t
        self._next_is_input = True
        return torch.empty(shape, dtype=dtype, device=device)

    def __call__(self, output_tensor, input_tensor, group, op, async_op: bool = False):
        world = dist.get_world_size(group)
        if world != self.symm.world:
            raise VB200Error("reduce-scatter group does not match the symmetric-memory group")
        if input_tensor.dtype != torch.float32:
            raise VB200Error(
                f"B200 reduce-scatter reduces in fp32 (VeOmni's default reduce_dtype, arguments_types.py:248-255); got {input_tensor.dtype}"
            )
        _RS_INPUTS.pop(input_tensor.data_ptr(), None)  # the copy-in / pre-divide hooks for this buffer have run
        pend = self._pending.pop(input_tensor.data_ptr(), None)
        if pend is not None:  # push mode: the gradients are still where autograd left them
            grads, plan, row = pend
            if row * world != input_tensor.numel() or output_tensor.numel() != row:
                raise VB200Error("reduce-scatter: gradient plan does not match the input / output tensors")
            desc = []
            for g, (numel, chunk, _off) in zip(grads, plan):
                desc += [g.data_ptr(), numel, chunk]
            with prof.span("fsdp_reduce_scatter", row * 2 * (world - 1)):
                self.symm.reduce_scatter_push_bf16(input_tensor, desc, row, output_tensor, _reduce_scale(op, world),
                                                   CH_REDUCE_SCATTER, self.num_ctas)
            cur = torch.cuda.current_stream()
            for g in grads:  # allocated by autograd on the compute stream, read by the kernel on this one
                g.record_stream(cur)
            return None
        chunk = self._packed.pop(input_tensor.data_ptr(), None)
        if chunk is not None:  # our copy-in left bf16 gradients in the first half of this buffer
            if chunk * world != input_tensor.numel():
                raise VB200Error("reduce-scatter: packed gradient size does not match the input tensor")
            self.symm.reduce_scatter_bf16(input_tensor, chunk, output_tensor, _reduce_scale(op, world), CH_REDUCE_SCATTER,
                                          self.num_ctas)
            return None
        if not self.symm.contains(input_tensor):  # allocate() order assumption broken: stay correct
            staged = self.symm.empty(tuple(input_tensor.shape), input_tensor.dtype, arena="fsdp_rs")
            staged.copy_(input_tensor)
            input_tensor = staged
        self.symm.reduce_scatter_f32(input_tensor, output_tensor, _reduce_scale(op, world), CH_REDUCE_SCATTER,
                                     self.num_ctas)
        return None


def pack_plan(shapes: list[tuple[int, ...]], world: int):
    """Geometry of torch._chunk_cat(grads, dim=0, num_chunks=world): per parameter (numel, chunk elements, row offset)
    and the row length S. Host-only (tests/test_parallel_host.py checks it against chunk_cat itself)."""
    plan, off = [], 0
    for shp in shapes:
        d0 = shp[0] if len(shp) else 1
        inner = 1
        for x in shp[1:]:
            inner *= x
        chunk = (d0 + world - 1) // world * inner
        plan.append((d0 * inner, chunk, off))
        off += chunk
    return plan, off


_orig_copy_in = None
_orig_div = None


def _copy_in(unsharded_grads, reduce_scatter_input, world_size):
    """Drop-in for torch's foreach_reduce_scatter_copy_in (_fsdp_collectives.py:667-675)."""
    comm = _RS_INPUTS.get(reduce_scatter_input.data_ptr())
    ok = (comm is not None and comm.pack_bf16 and world_size > 1 and reduce_scatter_input.dtype == torch.float32
          and all(g.is_cuda and g.dtype == torch.bfloat16 and g.is_contiguous() for g in unsharded_grads))
    if not ok:
        return _orig_copy_in(unsharded_grads, reduce_scatter_input, world_size)
    plan, row = pack_plan([tuple(g.shape) for g in unsharded_grads], world_size)
    if row * world_size != reduce_scatter_input.numel():
        return _orig_copy_in(unsharded_grads, reduce_scatter_input, world_size)
    if comm.mode == "push" and len(plan) <= 64:
        # nothing is copied: the reduce-scatter kernel reads the gradients in place. FSDP2 clears its own list right
        # after this call (:528-529); the references kept here keep the memory alive until the kernel is enqueued.
        comm._pending[reduce_scatter_input.data_ptr()] = (list(unsharded_grads), plan, row)
        return None
    import ctypes

    from . import _lib
    from ._lib import check, stream_ptr

    flat = []
    for g, (numel, chunk, off) in zip(unsharded_grads, plan):
        flat += [g.data_ptr(), numel, chunk, off]
    arr = (ctypes.c_int64 * len(flat))(*flat)
    with torch.cuda.device(reduce_scatter_input.device):
        check(_lib.load().vb200_fsdp_pack_bf16(arr, len(plan), world_size, row, reduce_scatter_input.data_ptr(), 0, stream_ptr()),
              "vb200_fsdp_pack_bf16")
    comm._packed[reduce_scatter_input.data_ptr()] = row
    return None


def _div_if_needed(tensor, div_factor):
    comm = _RS_INPUTS.get(tensor.data_ptr())
    if div_factor is not None and div_factor != 1 and comm is not None and (
            tensor.data_ptr() in comm._packed or tensor.data_ptr() in comm._pending):
        raise VB200Error("B200 reduce-scatter: a pre-divide factor on bf16-packed gradients is not supported "
                         "(use reduce_dtype=float32 without a custom divide factor, or pack_bf16=False)")
    return _orig_div(tensor, div_factor)


_ag_lib = None
_orig_get_inputs = None
# bf16 placeholder (address) -> the fp32 master shard it stands for, set by _param_all_gather_inputs and consumed by
# _ag_copy_in within the same foreach_all_gather call
_AG_SOURCES: dict[int, torch.Tensor] = {}


def _param_all_gather_inputs(fsdp_params):
    """Replacement for torch's _get_param_all_gather_inputs (_fsdp_collectives.py:295-343). The original casts the fp32
    master shards into a flat bf16 temporary (``_foreach_copy_``) which ``all_gather_copy_in`` then copies again into
    the all-gather buffer; here the flat temporary is only *allocated* — it gives FSDP the dtype / numel metadata it
    derives from the inputs — and the fp32 shards are remembered so :func:`_ag_copy_in` can cast them straight into the
    all-gather buffer (one pass of 4 + 2 bytes per parameter instead of 4 + 2 + 2 + 2)."""
    from torch.distributed.fsdp._fully_shard._fsdp_common import compiled_autograd_enabled
    from torch.distributed.fsdp._fully_shard._fsdp_param import ShardedState

    def fast(p) -> bool:
        return (p.param_dtype == torch.bfloat16 and not p.offload_to_cpu
                and not hasattr(p._sharded_local_tensor, "fsdp_pre_all_gather"))

    _AG_CALL.clear()
    if _ag_lib is None or compiled_autograd_enabled() or not all(fast(p) for p in fsdp_params):
        return _orig_get_inputs(fsdp_params)
    srcs = [p._sharded_param_data if p.sharded_state == ShardedState.SHARDED else p._sharded_post_forward_param_data
            for p in fsdp_params]
    if not all(t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16) for t in srcs):
        return _orig_get_inputs(fsdp_params)
    numels = [t.numel() for t in srcs]
    _AG_CALL[:] = list(fsdp_params)  # lets B200AllGather.__call__ deliver straight into these parameters
    flat = torch.empty((sum(numels),), device=srcs[0].device, dtype=torch.bfloat16)  # never written or read
    splits = torch.split(flat, numels)
    _AG_SOURCES.clear()
    for sp, t in zip(splits, srcs):
        if t.numel():
            _AG_SOURCES[sp.data_ptr()] = t
    return [[sp] for sp in splits]


def _ag_copy_in(all_gather_inputs, all_gather_output, inp_split_sizes, all_gather_input_numel, rank):
    """CUDA implementation of the ``fsdp::all_gather_copy_in`` op (all_gather_copy_in_cuda,
    _fsdp_collectives.py:175-188): the local shards (fp32 masters under mixed precision) written, cast to bf16, into
    this rank's slice of the all-gather buffer — one multi-tensor kernel instead of chunked ``_foreach_copy_``
    launches (measured 1.7 TB/s on Qwen3-8B). Anything but fp32/bf16 sources into a bf16 buffer takes torch's path."""
    all_gather_input = all_gather_output.narrow(0, all_gather_input_numel * rank, all_gather_input_numel)
    srcs = [_AG_SOURCES.pop(t.data_ptr(), t) if t.numel() else t for t in all_gather_inputs]
    _AG_SOURCES.clear()
    ok = (all_gather_output.is_cuda and all_gather_output.dtype == torch.bfloat16 and len(srcs) > 0
          and all(t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16) for t in srcs)
          and all(t.numel() == n for t, n in zip(srcs, inp_split_sizes)) and sum(inp_split_sizes) == all_gather_input_numel)
    if not ok:
        with torch.no_grad():
            torch._foreach_copy_(torch.split(all_gather_input, inp_split_sizes), srcs)
        return all_gather_input, all_gather_output
    import ctypes

    from . import _lib
    from ._lib import check, stream_ptr

    lib = _lib.load()
    for dt, code in ((torch.bfloat16, 0), (torch.float32, 1)):  # one launch per source dtype
        flat, off = [], 0
        for t in srcs:
            n = t.numel()
            if n and t.dtype == dt:
                flat += [t.data_ptr(), n, n, off]
            off += n
        if flat:
            arr = (ctypes.c_int64 * len(flat))(*flat)
            with torch.cuda.device(all_gather_output.device):
                check(lib.vb200_fsdp_pack_bf16(arr, len(flat) // 4, 1, all_gather_input_numel, all_gather_input.data_ptr(), code,
                                               stream_ptr()), "vb200_fsdp_pack_bf16")
    return all_gather_input, all_gather_output


def _patch_copy_in() -> None:
    """Route FSDP2's reduce-scatter copy-in through :func:`_copy_in`. ``foreach_reduce`` looks both names up in its
    module at call time, so replacing the module attributes is enough."""
    global _orig_copy_in, _orig_div
    from torch.distributed.fsdp._fully_shard import _fsdp_collectives as fc

    # private names of torch's FSDP2 (present in 2.9 .. 2.11): fail with a clear message instead of half-patching
    missing = [n for n in ("foreach_reduce_scatter_copy_in", "_div_if_needed", "_get_param_all_gather_inputs") if not hasattr(fc, n)]
    if missing:
        raise VB200Error(f"torch {torch.__version__}: FSDP2 internals {missing} not found in _fsdp_collectives; the fused "
                         "copy-in / copy-out patches cannot be installed (use install_fsdp_comm(..., rs_mode='f32', fuse_copy_out=False))")
    if _orig_copy_in is None:
        _orig_copy_in, _orig_div = fc.foreach_reduce_scatter_copy_in, fc._div_if_needed
        fc.foreach_reduce_scatter_copy_in = _copy_in
        fc._div_if_needed = _div_if_needed
        # the all-gather copy-in is a registered op: override its CUDA kernel
        global _ag_lib, _orig_get_inputs
        import warnings

        _ag_lib = torch.library.Library("fsdp", "IMPL")
        with warnings.catch_warnings():  # overriding a registered kernel is the point; torch warns about it once
            warnings.simplefilter("ignore")
            _ag_lib.impl("all_gather_copy_in", _ag_copy_in, "CUDA", allow_override=True)
            _ag_lib.impl("split_with_sizes_copy", _split_with_sizes_copy, "CUDA", allow_override=True)
        _orig_get_inputs = fc._get_param_all_gather_inputs
        fc._get_param_all_gather_inputs = _param_all_gather_inputs


def plan_fsdp_region(model: torch.nn.Module, world: int, param_bytes: int = 2, reduce_bytes: int = 4, modules=None,
                     misc_bytes: int = 256 << 20):
    """Size the symmetric region from the FSDP2 unit sizes of ``model``.

    Live all-gather outputs: the unit being copied out + the prefetched one; live reduce-scatter inputs:
    the one in flight + the previous one FSDP2 still holds. Returns (total_bytes, arena fractions).
    """
    from torch.distributed.fsdp import FSDPModule

    units = []
    for m in (model.modules() if modules is None else modules):
        if not isinstance(m, FSDPModule):
            continue
        nested = {id(p) for c in m.modules() if c is not m and isinstance(c, FSDPModule) for p in c.parameters()}
        numel = 0
        for p in m.parameters():
            if id(p) in nested:
                continue
            shape = tuple(p.shape)
            d0 = shape[0] if shape else 1
            rest = 1
            for s in shape[1:]:
                rest *= s
            numel += (d0 + world - 1) // world * world * rest
        units.append(numel)
    if not units:
        raise VB200Error("plan_fsdp_region: the model has no FSDP2 (fully_shard) modules")
    units.sort(reverse=True)
    top, second = units[0], (units[1] if len(units) > 1 else units[0])
    slack = 64 << 20
    ag = (top + 2 * second) * param_bytes + slack
    rs = (top + 2 * second) * reduce_bytes + slack
    misc = int(misc_bytes)  # everything that is not an FSDP buffer: Ulysses / EP staging, self-check scratch
    total = ag + rs + misc
    return total, {"fsdp_ag": ag / total, "fsdp_rs": rs / total, "misc": misc / total}


def install_fsdp_comm(model: torch.nn.Module, group: dist.ProcessGroup | None = None, symm: SymmetricMemory | None = None,
                      num_ctas: int = 32, pack_bf16: bool = True, rs_mode: str | None = None,
                      fuse_copy_out: bool = True, modules=None, misc_bytes: int = 256 << 20) -> SymmetricMemory:
    """Swap every FSDP2 unit of ``model`` onto the NVLink collectives.

    Call right after ``build_parallelize_model`` returns (veomni/trainer/base.py:387-404). ``rs_mode``: "push" (default:
    copy-in fused into the reduce-scatter kernel), "pull" (bf16 pack kernel + pull-reduce) or "f32" (torch's fp32
    copy-in + fp32 pull-reduce; what ``pack_bf16=False`` selects). ``fuse_copy_out``: the all-gather kernel stores into
    the per-parameter unsharded tensors directly. ``modules``: the FSDP2 units to switch (default: all of ``model``) — units
    sharded over another group (experts on ``ep_fsdp``) get their own call with that group.
    """
    from torch.distributed.fsdp import FSDPModule

    if symm is None:
        g = group if group is not None else dist.group.WORLD
        total, arenas = plan_fsdp_region(model, dist.get_world_size(g), modules=modules, misc_bytes=misc_bytes)
        symm = get_symmetric_memory(g, total, arenas)
    ag = B200AllGather(symm, num_ctas, fuse_copy_out)
    rs = B200ReduceScatter(symm, num_ctas, pack_bf16, rs_mode)
    _patch_copy_in()
    n = 0
    for m in (model.modules() if modules is None else modules):
        if isinstance(m, FSDPModule):
            m.set_custom_all_gather(ag)
            m.set_custom_reduce_scatter(rs)
            n += 1
    if n == 0:
        raise VB200Error("install_fsdp_comm: the model has no FSDP2 (fully_shard) modules")
    return symm


def uninstall_fsdp_comm(model: torch.nn.Module) -> list:
    """Put PyTorch's default NCCL collectives back on every FSDP2 unit (A/B measurements, debugging) and return what was
    installed, for :func:`reinstall_fsdp_comm`. The module-level hooks stay; they only act on buffers handed out by a B200
    comm object."""
    from torch.distributed.fsdp import FSDPModule
    from torch.distributed.fsdp._fully_shard._fsdp_collectives import DefaultAllGather, DefaultReduceScatter

    _AG_CALL.clear()
    saved = []
    for m in model.modules():
        if isinstance(m, FSDPModule):
            grp = m._get_fsdp_state()._fsdp_param_group
            if grp is None:
                continue
            saved.append((m, grp._all_gather_comm, grp._reduce_scatter_comm))
            m.set_custom_all_gather(DefaultAllGather())
            m.set_custom_reduce_scatter(DefaultReduceScatter())
    return saved


def reinstall_fsdp_comm(model: torch.nn.Module, saved: list) -> None:
    for m, ag, rs in saved:
        m.set_custom_all_gather(ag)
        m.set_custom_reduce_scatter(rs)
