"""Oracle for the collectives on the path: Ulysses all-to-all, FSDP2 all-gather / reduce-scatter.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  Functions take *lists indexed by rank*
and return lists indexed by rank; the exchange is done by slicing, which is the definition of the
collective.
"""

from __future__ import annotations

import torch


# ---- Ulysses ---------------------------------------------------------------------------------
def all_to_all_tensor(xs: list[torch.Tensor], scatter_dim: int, gather_dim: int) -> list[torch.Tensor]:
    """``all_to_all_tensor`` of every rank of the SP group.

    Reference: veomni/distributed/sequence_parallel/ulysses.py:125-135 (-> _all_to_all_single :86-122
    for dims <= 1, _all_to_all :64-83 otherwise).  Both paths compute: split the local tensor into
    P chunks along ``scatter_dim``, send chunk r to rank r, concatenate what arrives (in source-rank
    order) along ``gather_dim``.
    """
    P = len(xs)
    chunks = [torch.tensor_split(x, P, dim=scatter_dim) for x in xs]
    return [torch.cat([chunks[s][r] for s in range(P)], dim=gather_dim).contiguous() for r in range(P)]


def gather_seq_scatter_heads(xs, seq_dim: int, head_dim: int):
    """Reference: ulysses.py:235-253 (``_SeqAllToAll.apply(group, x, head_dim, seq_dim)``)."""
    return all_to_all_tensor(xs, scatter_dim=head_dim, gather_dim=seq_dim)


def gather_heads_scatter_seq(xs, head_dim: int, seq_dim: int):
    """Reference: ulysses.py:220-232 — pads the sequence dim with zeros to a multiple of P first."""
    P = len(xs)
    out = []
    for x in xs:
        n = x.size(seq_dim)
        if n % P:
            pad_shape = list(x.shape)
            pad_shape[seq_dim] = P - n % P
            x = torch.cat([x, torch.zeros(pad_shape, dtype=x.dtype)], dim=seq_dim)
        out.append(x)
    return all_to_all_tensor(out, scatter_dim=seq_dim, gather_dim=head_dim)


def all_to_all_rows(xs: list[torch.Tensor], splits: list[list[int]]) -> list[torch.Tensor]:
    """``_AlltoAllRegion.forward`` (veomni/distributed/sequence_parallel/ulysses.py:298-310) on every rank of a group at
    once: rank ``s`` splits its rows by ``splits[s]`` (``x.split(input_splits)``), ``dist.all_to_all`` hands block ``d`` to
    rank ``d``, and the receiver concatenates what it got in source-rank order. The reference's list-form all-to-all does
    not run on gloo, so this restatement of the definition is the checker: **parity unpinned** against an
    executed reference for this one function (it is pure indexing)."""
    world = len(xs)
    blocks = [list(x[: sum(sp)].split(list(sp), dim=0)) for x, sp in zip(xs, splits)]
    return [torch.cat([blocks[s][d] for s in range(world)], dim=0) for d in range(world)]


def repeat_kv_for_ulysses(key: torch.Tensor, ulysses_size: int) -> torch.Tensor:
    """KV head replication when P > Hkv.  Reference: veomni/ops/kernels/attention/__init__.py:245-255.
    key: [..., S, Hkv, D] with heads at dim -2."""
    kv = key.shape[-2]
    if ulysses_size > kv:
        return torch.repeat_interleave(key, dim=-2, repeats=ulysses_size // kv)
    return key


# ---- FSDP2 -----------------------------------------------------------------------------------
def fsdp_all_gather(shards: list[torch.Tensor], param_dtype: torch.dtype) -> torch.Tensor:
    """What every rank holds after one FSDP2 unit all-gather.

    Reference: torch/distributed/fsdp/_fully_shard/_fsdp_collectives.py:237-291 as configured by
    veomni/distributed/torch_parallelize.py:198-205 (bf16 params): each rank casts its flat fp32
    shard to ``param_dtype`` (copy-in :169-188) and the gathered buffer is the rank-order
    concatenation (all_gather_into_tensor :81-95).
    """
    return torch.cat([s.to(param_dtype) for s in shards], dim=0)


def fsdp_reduce_scatter(grads: list[torch.Tensor], reduce_dtype: torch.dtype, divide_factor: float | None):
    """Sharded gradient each rank holds after one FSDP2 unit reduce-scatter.

    Reference: _fsdp_collectives.py:448-660: unsharded grads are packed into a flat ``reduce_dtype``
    buffer (chunk_cat :220-234), summed across ranks, rank r keeps chunk r, then divided by the
    factor from _get_gradient_divide_factors (:701-759; fp32 reduce => SUM then post-divide by the
    world size, or by the factor VeOmni sets for EP, torch_parallelize.py:306-313).
    Summation order here is rank 0..N-1 in ``reduce_dtype`` (NCCL's ring order differs; the GPU
    kernel fixes the same 0..N-1 order, so the comparison is exact up to fp32 non-associativity of
    this fixed order, i.e. bitwise).
    """
    N = len(grads)
    flat = [g.to(reduce_dtype).reshape(N, -1) for g in grads]
    outs = []
    for r in range(N):
        acc = flat[0][r].clone()
        for s in range(1, N):
            acc = acc + flat[s][r]
        if divide_factor is not None:
            acc = acc / divide_factor
        outs.append(acc)
    return outs
