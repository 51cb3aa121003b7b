"""Ulysses sequence-parallel all-to-all on the NVLink pull kernel.

Mirrors veomni/distributed/sequence_parallel/ulysses.py: ``all_to_all_tensor`` (:125-135) is the
single choke point VeOmni's ``_SeqAllToAll`` (:138-160), ``gather_seq_scatter_heads`` (:235-253) and
``gather_heads_scatter_seq`` (:220-232) go through; :func:`install` swaps it for the kernel below, and
the functions here can also be called directly (same names, same argument meaning).

The reference implements the exchange as reshape/transpose/contiguous + ``all_to_all_single`` +
split/cat (:96-121): three extra HBM passes around an NCCL send/recv. Here the local tensor is staged
once into this rank's symmetric region and ONE kernel per call pulls, from every peer, exactly the
``[rows, seg]`` blocks this rank needs, writing them at their final position of the output tensor.
"""

from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist

from . import prof
from ._lib import VB200Error
from .symm import SymmetricMemory, get_symmetric_memory

CH_ULYSSES = 2
CH_IMAGES_SPLITS = 6
CH_IMAGES = 7


def a2a_plan(shape: tuple[int, ...], scatter_dim: int, gather_dim: int, world: int, itemsize: int):
    """Host-side geometry of the exchange for a contiguous local tensor of ``shape``.

    Returns (out_shape, desc) with desc = (src_rank_stride, src_row_stride, dst_peer_stride,
    dst_row_stride, rows, seg_bytes) in bytes, or raises for layouts the kernel does not cover.
    Supported: the two dims are adjacent (in either order) and everything before them has size 1 —
    i.e. packed ``[S, H, D]`` (dims 0/1) and ``[1, S, H, D]`` (dims 1/2), the layouts VeOmni's attention
    wrapper produces (veomni/ops/kernels/attention/__init__.py:257-281,322-330).
    """
    nd = len(shape)
    scatter_dim %= nd
    gather_dim %= nd
    lo, hi = min(scatter_dim, gather_dim), max(scatter_dim, gather_dim)
    if hi != lo + 1:
        raise VB200Error(f"ulysses a2a: scatter/gather dims must be adjacent, got {scatter_dim}/{gather_dim}")
    for d in shape[:lo]:
        if d != 1:
            raise VB200Error("ulysses a2a: leading dims before the exchanged pair must have size 1")
    inner = itemsize
    for d in shape[hi + 1:]:
        inner *= d
    A, B = shape[lo], shape[hi]  # local [A, B, inner]
    out = list(shape)
    if scatter_dim == hi:
        # gather dim lo (sequence), scatter dim hi (heads): local [Sl, H, *] -> out [P*Sl, H/P, *]
        if B % world:
            raise VB200Error(f"ulysses a2a: scattered dim {B} not divisible by the group size {world}")
        seg = (B // world) * inner
        desc = (seg, B * inner, A * seg, seg, A, seg)
        out[lo], out[hi] = A * world, B // world
    else:
        # scatter dim lo (sequence), gather dim hi (heads): local [S, Hl, *] -> out [S/P, P*Hl, *]
        if A % world:
            raise VB200Error(f"ulysses a2a: scattered dim {A} not divisible by the group size {world}")
        seg = B * inner
        rows = A // world
        desc = (rows * seg, seg, seg, world * seg, rows, seg)
        out[lo], out[hi] = rows, B * world
    if seg % 16:
        raise VB200Error(f"ulysses a2a: segment of {seg} bytes is not a multiple of 16")
    return tuple(out), desc


class _Stage:
    """Persistent staging buffers in the symmetric region, one per (slot, size)."""

    def __init__(self, symm: SymmetricMemory):
        self.symm = symm
        self.bufs: dict[tuple[int, int], torch.Tensor] = {}

    def get(self, slot: int, nbytes: int) -> torch.Tensor:
        key = (slot, nbytes)
        if key not in self.bufs:
            self.bufs[key] = self.symm.empty((nbytes,), torch.uint8, arena="misc")
        return self.bufs[key]

    def grow(self, tag: str, nbytes: int) -> torch.Tensor:
        """ONE buffer per tag, grown geometrically (payloads whose size changes from call to call: image rows). The old
        block returns to the arena; every collective publishes the offset of the buffer it exposes and is stream-ordered
        after the previous user, so ranks growing at different times is fine."""
        nbytes = (max(int(nbytes), 1) + 255) // 256 * 256
        cur = self.bufs.get((tag, 0))
        if cur is None or cur.numel() < nbytes:
            want = nbytes if cur is None else max(nbytes, cur.numel() * 3 // 2)
            self.bufs.pop((tag, 0), None)
            del cur
            self.bufs[(tag, 0)] = self.symm.empty(((want + 255) // 256 * 256,), torch.uint8, arena="misc")
        return self.bufs[(tag, 0)]


_stages: dict[int, _Stage] = {}


def _stage_for(symm: SymmetricMemory) -> _Stage:
    if id(symm) not in _stages:
        _stages[id(symm)] = _Stage(symm)
    return _stages[id(symm)]


def staging_views(shapes: list[tuple[int, ...]], dtype: torch.dtype, group: dist.ProcessGroup | None = None,
                  symm: SymmetricMemory | None = None) -> list[torch.Tensor]:
    """Views into the persistent send-staging block that the next ``all_to_all_many`` of the same shapes will use.
    A producer (e.g. the fused q/k-norm + RoPE kernel) can write its output there, which makes the exchange
    copy-free for those tensors ("optional fusion of RoPE+qk-RMSNorm on the way out")."""
    symm = symm if symm is not None else get_symmetric_memory(group)
    stage = _stage_for(symm)
    nbytes = [_numel(s) * dtype.itemsize for s in shapes]
    sizes = [(n + 255) // 256 * 256 for n in nbytes]
    buf = stage.get(len(shapes), sum(sizes))
    views, off = [], 0
    for shp, n, sz in zip(shapes, nbytes, sizes):
        views.append(buf[off : off + n].view(dtype).view(*shp))
        off += sz
    return views


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def all_to_all_many(xs: list[torch.Tensor], scatter_dim: int, gather_dim: int, group: dist.ProcessGroup | None = None,
                    symm: SymmetricMemory | None = None, num_ctas: int = 32) -> list[torch.Tensor]:
    """Exchange up to 4 tensors (e.g. q, k, v) in ONE kernel launch."""
    if not 1 <= len(xs) <= 4:
        raise VB200Error("all_to_all_many takes 1..4 tensors")
    symm = symm if symm is not None else get_symmetric_memory(group)
    world = symm.world
    if world == 1:
        return [x for x in xs]
    stage = _stage_for(symm)
    outs, descs, base = [], [], None
    # one staging block holding all inputs back to back (256-byte aligned pieces)
    sizes = [(x.numel() * x.element_size() + 255) // 256 * 256 for x in xs]
    buf = stage.get(len(xs), sum(sizes))
    off = 0
    for x, sz in zip(xs, sizes):
        if not x.is_cuda:
            raise VB200Error("ulysses a2a runs on CUDA tensors only (no gloo/CPU fallback)")
        out_shape, d = a2a_plan(tuple(x.shape), scatter_dim, gather_dim, world, x.element_size())
        n = x.numel() * x.element_size()
        slot = buf[off : off + n].view(x.dtype).view(x.shape)
        if x.data_ptr() != slot.data_ptr() or not x.is_contiguous():
            slot.copy_(x)  # stage (one local pass) unless the producer already wrote into the staging view
        out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
        descs.append((off, d[0], d[1], out.data_ptr(), d[2], d[3], d[4], d[5]))
        outs.append(out)
        off += sz
        base = buf
    moved = sum(x.numel() * x.element_size() for x in xs) * (world - 1) // world  # bytes this rank receives over NVLink
    with prof.span("ulysses_a2a", moved):
        symm.all_to_all(base, descs, CH_ULYSSES, num_ctas)
    return outs


def all_to_all_tensor(x: torch.Tensor, scatter_dim: int, gather_dim: int, group: dist.ProcessGroup | None = None,
                      async_op: bool = False):
    """Drop-in for veomni.distributed.sequence_parallel.ulysses.all_to_all_tensor (:125-135)."""
    out = all_to_all_many([x], scatter_dim, gather_dim, group)[0]
    if async_op:
        return lambda: out  # stream-ordered: nothing to wait for on the host
    return out


class _SeqAllToAll(torch.autograd.Function):
    """Same contract as the reference's _SeqAllToAll (ulysses.py:138-160): backward is the reverse exchange."""

    @staticmethod
    def forward(ctx: Any, group, local_input: torch.Tensor, scatter_dim: int, gather_dim: int) -> torch.Tensor:
        ctx.group, ctx.scatter_dim, ctx.gather_dim = group, scatter_dim, gather_dim
        return all_to_all_tensor(local_input, scatter_dim, gather_dim, group)

    @staticmethod
    def backward(ctx: Any, grad_output: torch.Tensor):
        return None, all_to_all_tensor(grad_output.contiguous(), ctx.gather_dim, ctx.scatter_dim, ctx.group), None, None


class _SeqAllToAllQKV(torch.autograd.Function):
    """q, k, v exchanged in one launch (forward) and their grads in one launch (backward)."""

    @staticmethod
    def forward(ctx: Any, group, scatter_dim: int, gather_dim: int, *xs: torch.Tensor):
        ctx.group, ctx.scatter_dim, ctx.gather_dim = group, scatter_dim, gather_dim
        return tuple(all_to_all_many(list(xs), scatter_dim, gather_dim, group))

    @staticmethod
    def backward(ctx: Any, *grads: torch.Tensor):
        gs = all_to_all_many([g.contiguous() for g in grads], ctx.gather_dim, ctx.scatter_dim, ctx.group)
        return (None, None, None, *gs)


def gather_seq_scatter_heads(x: torch.Tensor, seq_dim: int, head_dim: int, unpadded_dim_size: int = 0, group=None):
    """Reference: ulysses.py:235-253."""
    symm = get_symmetric_memory(group)
    x = _SeqAllToAll.apply(group, x, head_dim, seq_dim)
    if unpadded_dim_size and unpadded_dim_size % symm.world != 0:
        x = x.narrow(seq_dim, 0, unpadded_dim_size)
    return x


def gather_seq_scatter_heads_qkv(q, k, v, seq_dim: int, head_dim: int, group=None):
    """q/k/v variant: one kernel launch for the three tensors."""
    return _SeqAllToAllQKV.apply(group, head_dim, seq_dim, q, k, v)


def gather_heads_scatter_seq(x: torch.Tensor, head_dim: int, seq_dim: int, group=None):
    """Reference: ulysses.py:220-232 (zero-pads the sequence dim to a multiple of the group size)."""
    symm = get_symmetric_memory(group)
    n = x.size(seq_dim)
    if n % symm.world:
        pad = list(x.shape)
        pad[seq_dim] = symm.world - n % symm.world
        x = torch.cat([x, x.new_zeros(pad)], dim=seq_dim)
    return _SeqAllToAll.apply(group, x, seq_dim, head_dim)


def images_chunks(splits: torch.Tensor, rank: int, row_bytes: int) -> torch.Tensor:
    """Block list of one uneven row exchange. ``splits[s, d]`` (integer tensor, host or device) = rows rank ``s`` sends to
    rank ``d``; returns int64 ``[world, 4]`` = {src_off, dst_off, bytes, peer} as the C struct ``ChunkDesc`` (csrc/p2p.cu)
    for ``rank``'s pull: the block from source ``s`` sits in ``s``'s send buffer behind the rows ``s`` sends to lower
    ranks, and lands behind the blocks of lower sources (``torch.cat(output_tensor_list)`` order, ulysses.py:309)."""
    splits = splits.to(torch.int64)
    world = splits.shape[0]
    excl = torch.cumsum(splits, dim=1) - splits
    n = splits[:, rank]
    dst = torch.cumsum(n, dim=0) - n
    peer = torch.arange(world, dtype=torch.int64, device=splits.device)
    return torch.stack([excl[:, rank] * row_bytes, dst * row_bytes, n * row_bytes, peer], dim=1).contiguous()


def _pull_rows(symm: SymmetricMemory, x: torch.Tensor, splits_dev: torch.Tensor, n_out: int, num_ctas: int) -> torch.Tensor:
    """Stage ``x`` (this rank's rows, grouped by destination) and pull the blocks addressed to this rank."""
    from . import _lib
    from ._lib import check, stream_ptr

    row_bytes = x[0].numel() * x.element_size() if x.shape[0] else _numel(x.shape[1:]) * x.element_size()
    if row_bytes % 16:
        raise VB200Error(f"all_to_all_images: rows of {row_bytes} bytes are not a multiple of 16")
    stage = _stage_for(symm)
    buf = stage.grow("images", x.numel() * x.element_size())
    if x.numel():
        buf[: x.numel() * x.element_size()].view(x.dtype).view(x.shape).copy_(x)
    out = torch.empty((max(n_out, 1), *x.shape[1:]), dtype=x.dtype, device=x.device)[:n_out]
    chunks = images_chunks(splits_dev, symm.rank, row_bytes)
    with torch.cuda.device(x.device), prof.span("images_a2a", out.numel() * out.element_size()):
        check(_lib.load().vb200_chunk_pull(symm.comm, CH_IMAGES, symm.offset_of(buf), chunks.data_ptr(), symm.world,
                                           out.data_ptr(), num_ctas, stream_ptr()), "vb200_chunk_pull")
    return out


class _AlltoAllRegion(torch.autograd.Function):
    """Same contract as the reference's ``_AlltoAllRegion`` (ulysses.py:298-316): uneven row exchange along dim 0,
    backward = the exchange with the split lists swapped. The reference issues a list-form ``dist.all_to_all``; here the
    ``[world, world]`` split matrix is all-gathered through the symmetric region (``world`` ints per rank, no host
    read-back), the block list is computed from it on the device, and ONE pull kernel moves the rows."""

    @staticmethod
    def forward(ctx: Any, group, x: torch.Tensor, input_splits, output_splits, num_ctas: int = 16, symm=None):
        symm = symm if symm is not None else get_symmetric_memory(group)
        if not x.is_cuda:
            raise VB200Error("all_to_all_images runs on CUDA tensors only (no gloo/CPU fallback)")
        if len(input_splits) != symm.world or len(output_splits) != symm.world or sum(input_splits) != x.shape[0]:
            raise VB200Error("all_to_all_images: split lists must have one entry per rank and cover the rows of x")
        w = symm.world
        mat = _stage_for(symm).grow("images_splits", w * w * 4).view(torch.int32)[: w * w]
        mat[symm.rank * w : (symm.rank + 1) * w].copy_(torch.tensor(list(input_splits), dtype=torch.int32), non_blocking=True)
        symm.all_gather_inplace(mat, w, CH_IMAGES_SPLITS, 1)
        splits = mat.view(w, w).to(torch.int64)  # a copy: the staging matrix is reused by the next call
        ctx.symm, ctx.num_ctas, ctx.n_in = symm, num_ctas, int(x.shape[0])
        ctx.save_for_backward(splits)
        return _pull_rows(symm, x.contiguous(), splits, int(sum(output_splits)), num_ctas)

    @staticmethod
    def backward(ctx: Any, dy: torch.Tensor):
        (splits,) = ctx.saved_tensors
        dx = _pull_rows(ctx.symm, dy.contiguous(), splits.t().contiguous(), ctx.n_in, ctx.num_ctas)
        return None, dx, None, None, None, None


def all_to_all_images(image_embeds: torch.Tensor, in_splits, out_splits, group: dist.ProcessGroup | None = None,
                      symm: SymmetricMemory | None = None):
    """Drop-in for veomni.distributed.sequence_parallel.ulysses.all_to_all_images (:319-324): balance the vision tower's
    image embeddings over the Ulysses group (``in_splits[d]`` rows go to rank ``d``, ``out_splits[s]`` rows arrive from
    rank ``s``); an empty ``in_splits`` is the identity, rows beyond ``sum(in_splits)`` are dropped."""
    if not in_splits:
        return image_embeds
    image_embeds = image_embeds[: sum(in_splits)]
    if group is None:
        from .parallel_state import get_parallel_state

        ps = get_parallel_state()
        if not ps.ulysses_enabled:
            return image_embeds
        group = ps.ulysses_group
    if dist.get_world_size(group) == 1:
        return image_embeds
    return _AlltoAllRegion.apply(group, image_embeds, list(in_splits), list(out_splits), 16, symm)


def install() -> None:
    """Route VeOmni's Ulysses exchanges through this module (no-op if VeOmni is not importable)."""
    try:
        from veomni.distributed import sequence_parallel as pkg
        from veomni.distributed.sequence_parallel import ulysses as ref
    except ImportError:
        return
    ref.all_to_all_tensor = all_to_all_tensor
    ref.all_to_all_images = pkg.all_to_all_images = all_to_all_images
