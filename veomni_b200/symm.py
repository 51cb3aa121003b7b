"""Symmetric (peer-mapped) device memory over NVLink/NVSwitch for one process group.

Each rank ``cudaMalloc``s one region ``[signal pad | data]`` through the C ABI, exports it with
CUDA IPC, and maps every peer's region (handles are exchanged once through ``torch.distributed``
— plumbing only).  :class:`SymmetricMemory` then hands out torch tensors that alias the local data
region (so FSDP2 / autograd can own them like any tensor) and owns the ``vb200_comm`` handle the
collective kernels take.  Offsets need not match across ranks: every collective publishes its local
offset to the peers together with its ready flag.

Reference role: this replaces ``torch.distributed``'s NCCL communicator for the collectives listed in
SURVEY.md §2.3 K1/K2/K4/K11; NCCL stays for scalar all-reduces (clip-grad-norm, loss).
"""

from __future__ import annotations

import ctypes
import os
import weakref
from ctypes import c_void_p

import torch
import torch.distributed as dist

from . import _lib
from ._lib import VB200Error, check

_ALIGN = 256


class Arena:
    """First-fit free list with coalescing over [base, base + size). Pure host bookkeeping."""

    def __init__(self, base: int, size: int):
        self.base, self.size = base, size
        self.free: list[tuple[int, int]] = [(base, size)]
        self.live = 0

    def alloc(self, nbytes: int) -> tuple[int, int]:
        size = max(_ALIGN, (int(nbytes) + _ALIGN - 1) // _ALIGN * _ALIGN)
        for i, (o, s) in enumerate(self.free):
            if s >= size:
                if s == size:
                    self.free.pop(i)
                else:
                    self.free[i] = (o + size, s - size)
                self.live += size
                return o, size
        raise VB200Error(f"symmetric arena exhausted: need {size} B of {self.size} B (live {self.live} B); "
                         "raise the region size passed to SymmetricMemory")

    def release(self, off: int, size: int) -> None:
        self.live -= size
        fl = sorted(self.free + [(off, size)])
        merged = [fl[0]]
        for o, s in fl[1:]:
            lo, ls = merged[-1]
            if lo + ls == o:
                merged[-1] = (lo, ls + s)
            else:
                merged.append((o, s))
        self.free = merged


class _Block:
    """Lifetime anchor of one allocation: torch keeps it alive as long as the storage lives."""

    def __init__(self, ptr: int, nbytes: int):
        self._ptr, self._n = ptr, nbytes

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self._n,), "typestr": "|u1", "data": (self._ptr, False), "version": 3, "strides": None}


class SymmetricMemory:
    """Peer-mapped region + comm handle for ``group`` (world size <= 8, one node)."""

    def __init__(self, group: dist.ProcessGroup | None, data_bytes: int, device: torch.device | None = None,
                 arenas: dict[str, float] | None = None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise VB200Error("SymmetricMemory supports up to 8 ranks (one NVSwitch domain)")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self._lib = lib
        self.pad_bytes = int(lib.vb200_comm_signal_bytes())
        self.data_bytes = (int(data_bytes) + _ALIGN - 1) // _ALIGN * _ALIGN
        total = self.pad_bytes + self.data_bytes
        with torch.cuda.device(self.device):
            base = c_void_p()
            check(lib.vb200_symm_alloc(ctypes.byref(base), total), "vb200_symm_alloc")
            self._base = base.value
            handle = ctypes.create_string_buffer(64)
            check(lib.vb200_ipc_get_handle(self._base, handle), "vb200_ipc_get_handle")
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (handle.raw, os.getpid()), group=self.group)
            self._peer_bases = []
            for p, (h, _pid) in enumerate(gathered):
                if p == self.rank:
                    self._peer_bases.append(self._base)
                else:
                    ptr = c_void_p()
                    check(lib.vb200_ipc_open_handle(ctypes.create_string_buffer(h, 64), ctypes.byref(ptr)),
                          "vb200_ipc_open_handle")
                    self._peer_bases.append(ptr.value)
            sig = (c_void_p * self.world)(*self._peer_bases)
            data = (c_void_p * self.world)(*[b + self.pad_bytes for b in self._peer_bases])
            comm = c_void_p()
            check(lib.vb200_comm_create(ctypes.byref(comm), self.rank, self.world, data, sig, self.data_bytes),
                  "vb200_comm_create")
            self.comm = comm.value
        self.data_ptr = self._base + self.pad_bytes
        # Arenas: independent first-fit free lists over disjoint slices of the data region. A block is
        # only ever reused by the same class of user (FSDP all-gather outputs / reduce-scatter inputs /
        # everything else), whose own stream protocol makes same-class reuse safe.
        self._arenas: dict[str, Arena] = {}
        arenas = arenas or {"misc": 1.0}
        tot = float(sum(arenas.values()))
        cur = 0
        names = list(arenas)
        for i, name in enumerate(names):
            size = int(self.data_bytes * (arenas[name] / tot)) // _ALIGN * _ALIGN
            if i == len(names) - 1:
                size = self.data_bytes - cur
            self._arenas[name] = Arena(cur, size)
            cur += size
        dist.barrier(group=self.group)  # every peer has mapped every region before first use

    # ---- allocation ---------------------------------------------------------------------------
    def alloc(self, nbytes: int, arena: str = "misc") -> tuple[int, torch.Tensor]:
        """Return (offset, uint8 tensor) of a block in the local data region; freed with the tensor."""
        if arena not in self._arenas:
            arena = "misc" if "misc" in self._arenas else next(iter(self._arenas))
        ar = self._arenas[arena]
        o, size = ar.alloc(nbytes)
        block = _Block(self.data_ptr + o, int(nbytes))
        weakref.finalize(block, ar.release, o, size)
        t = torch.as_tensor(block, device=self.device)
        return o, t

    def empty(self, shape, dtype: torch.dtype, arena: str = "misc") -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        _, t = self.alloc(max(1, n * dtype.itemsize), arena)
        return t[: n * dtype.itemsize].view(dtype).view(*shape)

    def contains(self, t: torch.Tensor) -> bool:
        off = t.data_ptr() - self.data_ptr
        return 0 <= off and off + t.numel() * t.element_size() <= self.data_bytes

    def offset_of(self, t: torch.Tensor) -> int:
        if not self.contains(t):
            raise VB200Error("tensor does not live in this symmetric region")
        return t.data_ptr() - self.data_ptr

    # ---- collectives (thin ctypes faces; all on the current stream) ----------------------------
    def barrier(self, channel: int = 31) -> None:
        with torch.cuda.device(self.device):
            check(self._lib.vb200_comm_barrier(self.comm, channel, _lib.stream_ptr()), "vb200_comm_barrier")

    def check(self) -> None:
        with torch.cuda.device(self.device):
            check(self._lib.vb200_comm_check(self.comm), "vb200_comm_check")

    def all_gather_inplace(self, out: torch.Tensor, shard_numel: int, channel: int, num_ctas: int = 32) -> None:
        """``out`` (in this region) holds N shards of ``shard_numel`` elements; slot ``rank`` is filled."""
        with torch.cuda.device(self.device):
            check(self._lib.vb200_allgather(self.comm, channel, self.offset_of(out), shard_numel * out.element_size(),
                                            num_ctas, _lib.stream_ptr()), "vb200_allgather")

    def all_gather_scatter(self, out: torch.Tensor, shard_numel: int, table: list[int], channel: int,
                           num_ctas: int = 32) -> None:
        """All-gather with the copy-out fused in. ``out`` (in this region) only lends its slot ``rank`` (this rank's
        shard row); ``table`` = n x (byte offset inside a shard row, shard bytes, destination pointer)."""
        arr = (ctypes.c_int64 * len(table))(*[int(v) for v in table])
        with torch.cuda.device(self.device):
            check(self._lib.vb200_allgather_scatter(self.comm, channel, self.offset_of(out), shard_numel * out.element_size(),
                                                    arr, len(table) // 3, num_ctas, _lib.stream_ptr()), "vb200_allgather_scatter")

    def reduce_scatter_push_bf16(self, staging: torch.Tensor, desc: list[int], row: int, out: torch.Tensor, scale: float,
                                 channel: int, num_ctas: int = 32) -> None:
        """Reduce-scatter with the copy-in fused in. ``staging`` (in this region, >= world*row*2 bytes) receives the
        peers' bf16 chunks; ``desc`` = n x (gradient pointer, numel, chunk elements)."""
        if staging.numel() * staging.element_size() < self.world * row * 2:
            raise VB200Error("reduce_scatter_push_bf16: staging buffer too small")
        arr = (ctypes.c_int64 * len(desc))(*[int(v) for v in desc])
        with torch.cuda.device(self.device):
            check(self._lib.vb200_reduce_scatter_push_bf16(self.comm, channel, self.offset_of(staging), arr, len(desc) // 3,
                                                           int(row), float(scale), out.data_ptr(), num_ctas,
                                                           _lib.stream_ptr()), "vb200_reduce_scatter_push_bf16")

    def reduce_scatter_f32(self, inp: torch.Tensor, out: torch.Tensor, scale: float, channel: int,
                           num_ctas: int = 32) -> None:
        chunk = inp.numel() // self.world
        with torch.cuda.device(self.device):
            check(self._lib.vb200_reduce_scatter_f32(self.comm, channel, self.offset_of(inp), chunk, float(scale),
                                                     out.data_ptr(), num_ctas, _lib.stream_ptr()),
                  "vb200_reduce_scatter_f32")

    def reduce_scatter_bf16(self, inp: torch.Tensor, chunk: int, out: torch.Tensor, scale: float, channel: int,
                            num_ctas: int = 32) -> None:
        """``inp``: a tensor of this region whose first ``world * chunk`` bf16 elements are the packed gradients."""
        with torch.cuda.device(self.device):
            check(self._lib.vb200_reduce_scatter_bf16(self.comm, channel, self.offset_of(inp), chunk, float(scale),
                                                      out.data_ptr(), num_ctas, _lib.stream_ptr()),
                  "vb200_reduce_scatter_bf16")

    def all_to_all(self, base: torch.Tensor, descs: list[tuple], channel: int, num_ctas: int = 32) -> None:
        """descs: (src_off, src_rank_stride, src_row_stride, dst_ptr, dst_peer_stride, dst_row_stride, rows, seg_bytes)
        with src offsets relative to ``base`` (a tensor in this region)."""
        flat = [int(v) for d in descs for v in d]
        arr = (ctypes.c_int64 * len(flat))(*flat)
        with torch.cuda.device(self.device):
            check(self._lib.vb200_all_to_all(self.comm, channel, self.offset_of(base), len(descs), arr, num_ctas,
                                             _lib.stream_ptr()), "vb200_all_to_all")

    def close(self) -> None:
        if getattr(self, "comm", None):
            torch.cuda.synchronize(self.device)
            self._lib.vb200_comm_destroy(self.comm)
            self.comm = None
            for p, b in enumerate(self._peer_bases):
                if p != self.rank:
                    self._lib.vb200_ipc_close_handle(b)
            self._lib.vb200_symm_free(self._base)


_default: dict[tuple, SymmetricMemory] = {}


def get_symmetric_memory(group: dist.ProcessGroup | None = None, data_bytes: int | None = None,
                         arenas: dict[str, float] | None = None, tag: str = "") -> SymmetricMemory:
    """One shared region per (process group, tag), created on first use. Users with their own sizing (the EP staging
    buffers) pass a ``tag`` so that they do not land in a region another user sized for itself — PyTorch hands out the
    same group object for equal rank sets, e.g. the FSDP shard group and the EP group of an all-EP job."""
    g = group if group is not None else dist.group.WORLD
    key = (id(g), tag)
    if key not in _default:
        if data_bytes is None:
            data_bytes = int(os.environ.get("VB200_SYMM_BYTES", str(1 << 30)))
        _default[key] = SymmetricMemory(g, data_bytes, arenas=arenas)
    return _default[key]
