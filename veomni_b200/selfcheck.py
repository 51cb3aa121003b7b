"""In-process multi-GPU parity checks of the NVLink collectives (run by bench.py at world > 1 and by
tests/multigpu_worker.py).

Every check moves the *expected* data with NCCL (``dist.all_gather`` of the inputs) and then applies the
definition of the collective with plain tensor indexing — rank-order concatenation for the all-gather, a fixed
rank-order fp32 sum for the reduce-scatter (the order ``oracle/comm.py`` pins against the reference), head / sequence
slicing for the Ulysses exchange (veomni/distributed/sequence_parallel/ulysses.py:64-122), expert-major /
source-minor ordering for the EP dispatch (veomni/distributed/moe/moe_layer.py:72-99) — and compares bit for bit.
The FSDP2 check trains the same small model with the custom comm and with PyTorch's NCCL comm.

``run_all`` returns ``{"allgather": "bit-exact", ...}`` and raises ``VB200Error`` on the first mismatch, so a bench line
carrying the dict certifies the collectives of that very process group on that very box.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import VB200Error
from .symm import SymmetricMemory

BF = torch.bfloat16


def _gather_all(x: torch.Tensor, group=None) -> list[torch.Tensor]:
    out = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, x.contiguous(), group=group)
    return out


def _fail(what: str):
    raise VB200Error(f"multi-GPU parity check failed: {what}")


def _rank_order_sum(xs: list[torch.Tensor], scale: float) -> torch.Tensor:
    acc = xs[0].float().clone()
    for t in xs[1:]:
        acc += t.float()
    return acc * scale


def check_allgather(symm: SymmetricMemory, dev, sizes=(8, 1000, 4096 * 3 + 2, 1 << 20)) -> str:
    rank, world = symm.rank, symm.world
    for numel in sizes:
        for dtype in (BF, torch.float32):
            out = symm.empty((numel * world,), dtype, arena="fsdp_ag")
            out.fill_(-1)
            g = torch.Generator(device="cpu").manual_seed(1000 * rank + numel)
            shard = torch.randn(numel, generator=g).to(dtype).to(dev)
            out[rank * numel : (rank + 1) * numel].copy_(shard)
            symm.all_gather_inplace(out, numel, 0)
            if not torch.equal(out, torch.cat(_gather_all(shard, symm.group))):
                _fail(f"allgather numel={numel} dtype={dtype}")
    symm.check()
    return "bit-exact"


def check_allgather_scatter(symm: SymmetricMemory, dev) -> str:
    """The fused copy-out: per-parameter destinations [world * n_i] hold rank p's shard of parameter i at p * n_i."""
    rank, world = symm.rank, symm.world
    cases = [[4096 * 512, 4096, 128, 16, 1024 * 384], [10, 3, 28, 128, 5], [8] * 40]
    for ci, numels in enumerate(cases):
        for dtype in (BF, torch.float32):
            es = dtype.itemsize
            row = sum(numels)
            buf = symm.empty((row * world,), dtype, arena="fsdp_ag")
            buf.fill_(-1)
            g = torch.Generator(device="cpu").manual_seed(31 * rank + ci)
            shard = torch.randn(row, generator=g).to(dtype).to(dev)
            buf[rank * row : (rank + 1) * row].copy_(shard)
            dsts = [torch.full((n * world,), -2.0, dtype=dtype, device=dev) for n in numels]
            table, off = [], 0
            for n, d in zip(numels, dsts):
                table += [off * es, n * es, d.data_ptr()]
                off += n
            symm.all_gather_scatter(buf, row, table, 0)
            full = torch.stack(_gather_all(shard, symm.group))  # [world, row]
            off = 0
            for n, d in zip(numels, dsts):
                if not torch.equal(d.view(world, n), full[:, off : off + n]):
                    _fail(f"allgather+copy-out case {ci} dtype={dtype} param numel={n}")
                off += n
    symm.check()
    return "bit-exact"


def check_reduce_scatter(symm: SymmetricMemory, dev, chunks=(1, 7, 1024, 4099, 1 << 18)) -> dict:
    rank, world = symm.rank, symm.world
    scale = 1.0 / world
    for chunk in chunks:
        g = torch.Generator(device="cpu").manual_seed(77 * rank + chunk)
        x = torch.randn(chunk * world, generator=g).to(dev)
        inp = symm.empty((chunk * world,), torch.float32, arena="fsdp_rs")
        inp.copy_(x)
        out = torch.empty(chunk, dtype=torch.float32, device=dev)
        symm.reduce_scatter_f32(inp, out, scale, 1)
        ref = _rank_order_sum([t[rank * chunk : (rank + 1) * chunk] for t in _gather_all(x, symm.group)], scale)
        if not torch.equal(out, ref):
            _fail(f"reduce_scatter fp32 chunk={chunk}")
        xb = x.to(BF)
        buf = symm.empty((chunk * world,), torch.float32, arena="fsdp_rs")
        buf.view(BF)[: chunk * world].copy_(xb)
        outb = torch.empty(chunk, dtype=torch.float32, device=dev)
        symm.reduce_scatter_bf16(buf, chunk, outb, scale, 1)
        refb = _rank_order_sum([t[rank * chunk : (rank + 1) * chunk] for t in _gather_all(xb, symm.group)], scale)
        if not torch.equal(outb, refb):
            _fail(f"reduce_scatter bf16-pull chunk={chunk}")
        del inp, buf
    symm.check()
    return {"reduce_scatter_f32": "bit-exact", "reduce_scatter_bf16_pull": "bit-exact"}


def check_reduce_scatter_push(symm: SymmetricMemory, dev) -> str:
    """The fused copy-in: gradients read in place, chunk_cat layout (dim-0 zero padding), rank-order fp32 sum."""
    from .fsdp_comm import pack_plan

    rank, world = symm.rank, symm.world
    scale = 1.0 / world
    cases = [
        [(4096, 512), (4096,), (1024, 128), (128,), (512, 1024)],  # aligned: vector path
        [(10, 4), (3,), (7, 2, 2), (16, 8), (1, 5)],               # ragged: scalar path + dim-0 padding
        [(64, 8)] * 40,
    ]
    for ci, shapes in enumerate(cases):
        g = torch.Generator(device="cpu").manual_seed(13 * rank + ci)
        grads = [torch.randn(*s, generator=g).to(BF).to(dev) for s in shapes]
        plan, row = pack_plan(shapes, world)
        staging = symm.empty((row * world,), torch.float32, arena="fsdp_rs")
        staging.fill_(float("nan"))
        out = torch.full((row,), float("nan"), dtype=torch.float32, device=dev)
        desc = []
        for t, (numel, chunk, _off) in zip(grads, plan):
            desc += [t.data_ptr(), numel, chunk]
        for _ in range(2):  # twice: the second call reuses the staging buffer
            symm.reduce_scatter_push_bf16(staging, desc, row, out, scale, 1)
        # expected: chunk_cat of every rank's gradients, then the rank-order sum of this rank's row
        per_src = list(zip(*[_gather_all(t, symm.group) for t in grads]))  # [src][param]
        rows = []
        for src in range(world):
            packed = torch.empty(world, row, dtype=BF, device=dev)
            torch._chunk_cat(list(per_src[src]), dim=0, num_chunks=world, out=packed)
            rows.append(packed[rank].clone())
        if not torch.equal(out, _rank_order_sum(rows, scale)):
            _fail(f"reduce_scatter push (fused copy-in) case {ci}")
        del staging
    symm.check()
    return "bit-exact"


def check_ulysses(symm: SymmetricMemory, dev) -> str:
    from . import ulysses as U

    rank, world = symm.rank, symm.world
    for (Sl, H, D) in ((6, 4 * world, 8), (128, 8 * world, 128), (33, 2 * world, 64)):
        g = torch.Generator(device="cpu").manual_seed(5 * rank + Sl)
        x = torch.randn(Sl, H, D, generator=g).to(BF).to(dev)
        y = U.all_to_all_many([x], 1, 0, group=symm.group, symm=symm)[0]  # gather sequence, scatter heads
        xs = _gather_all(x, symm.group)
        hl = H // world
        ref = torch.cat([t[:, rank * hl : (rank + 1) * hl] for t in xs], dim=0)
        if not torch.equal(y, ref):
            _fail(f"ulysses gather_seq_scatter_heads {(Sl, H, D)}")
        z = U.all_to_all_many([y], 0, 1, group=symm.group, symm=symm)[0]  # back
        if not torch.equal(z, x):
            _fail(f"ulysses round trip {(Sl, H, D)}")
    symm.check()
    return "bit-exact"


def check_images(symm: SymmetricMemory, dev) -> str:
    """Uneven row exchange (``all_to_all_images``, ulysses.py:298-324): output == the blocks addressed to this rank in
    source order; gradient == the blocks of the peers' output gradients that came from this rank. Zero-row blocks included."""
    from . import ulysses as U

    rank, world = symm.rank, symm.world
    for trial, H in ((0, 64), (1, 1152)):
        splits = [[(3 * s + 5 * d + trial) % 7 * (1 + 40 * trial) for d in range(world)] for s in range(world)]
        ins, outs = splits[rank], [splits[s][rank] for s in range(world)]
        g = torch.Generator(device="cpu").manual_seed(11 * rank + trial)
        x = torch.randn(sum(ins) + 3, H, generator=g).to(BF).to(dev).requires_grad_(True)  # 3 trailing rows are dropped
        y = U.all_to_all_images(x, ins, outs, group=symm.group, symm=symm)
        cap = max(sum(r) for r in splits) + 3
        pad = torch.zeros(cap, H, dtype=BF, device=dev)
        pad[: x.shape[0]] = x.detach()
        xs = _gather_all(pad, symm.group)
        ref = torch.cat([xs[s][sum(splits[s][:rank]) : sum(splits[s][:rank]) + splits[s][rank]] for s in range(world)], dim=0)
        if y.shape != ref.shape or not torch.equal(y.detach(), ref):
            _fail(f"all_to_all_images forward, trial {trial}")
        dy = torch.randn(y.shape, generator=g).to(BF).to(dev)
        y.backward(dy)
        capo = max(sum(splits[s][d] for s in range(world)) for d in range(world))
        pad = torch.zeros(max(capo, 1), H, dtype=BF, device=dev)
        pad[: dy.shape[0]] = dy
        dys = _gather_all(pad, symm.group)
        parts = []
        for d in range(world):
            o = sum(splits[s][d] for s in range(rank))
            parts.append(dys[d][o : o + splits[rank][d]])
        ref_dx = torch.cat(parts + [torch.zeros(3, H, dtype=BF, device=dev)], dim=0)
        if not torch.equal(x.grad, ref_dx):
            _fail(f"all_to_all_images backward, trial {trial}")
    symm.check()
    return "bit-exact"


def check_ep_dispatch(symm: SymmetricMemory, dev) -> str:
    """Dispatched tokens == expert-major / source-minor / token-order selection of the gathered inputs; combine with
    unit weights returns K * hidden (exact in bf16 for K a power of two)."""
    from . import ep as EP

    rank, world = symm.rank, symm.world
    ctx = EP.EPContext(group=symm.group, symm=symm)
    E, K, H, T = 4 * world, 2, 64, 96
    g = torch.Generator(device="cpu").manual_seed(300 + rank)
    hs = torch.randn(T, H, generator=g).to(BF).to(dev)
    idx = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)]).to(dev)
    tokens, plan = EP.ep_dispatch(ctx, hs, idx, E)
    all_hs, all_idx = _gather_all(hs, symm.group), _gather_all(idx, symm.group)
    el = E // world
    exp = []
    for le in range(el):
        e = rank * el + le
        for s in range(world):
            flat = all_idx[s].flatten()  # (t, k) order == stable argsort order inside one expert
            sel = (flat == e).nonzero().flatten() // K
            exp.append(all_hs[s][sel])
    exp = torch.cat(exp) if exp else hs[:0]
    if tokens.shape != exp.shape or not torch.equal(tokens, exp):
        _fail("EP dispatch (token order / content)")
    w = torch.ones(T, K, dtype=BF, device=dev)
    back = EP.ep_combine(tokens, w, plan)
    if not torch.equal(back, (hs.float() * K).to(BF)):
        _fail("EP combine round trip")
    # routing collapse: every token picks rank 0's first two experts, so every other rank receives NO row and still has
    # to take part in all four exchanges (forward and backward of dispatch and combine)
    hs2 = hs.clone().requires_grad_(True)
    idx0 = torch.tensor([0, 1], device=dev).expand(T, K).contiguous()
    tok0, plan0 = EP.ep_dispatch(ctx, hs2, idx0, E)
    if (rank == 0) != (tok0.shape[0] == world * T * K) or (rank != 0 and tok0.shape[0] != 0):
        _fail("EP dispatch under routing collapse (row count)")
    back0 = EP.ep_combine(tok0, w, plan0)
    if not torch.equal(back0.detach(), (hs.float() * K).to(BF)):
        _fail("EP combine round trip under routing collapse")
    back0.float().sum().backward()
    if not torch.equal(hs2.grad, torch.full_like(hs, float(K))):
        _fail("EP backward under routing collapse")
    symm.check()
    return "bit-exact"


def check_fsdp(dev, world: int, group=None) -> dict:
    """Same toy FSDP2 model, same data: gradients with the B200 comm (every mode) vs PyTorch's NCCL comm, plus the
    EP-style gradient divide factor (veomni/distributed/torch_parallelize.py:306-313)."""
    from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

    from .fsdp_comm import install_fsdp_comm

    rank = dist.get_rank(group)

    def build(reshard=True, factor=None):
        torch.manual_seed(3)
        dims = [256, 256, 250, 130, 256]  # ragged middle layers: dim-0 padding and the scalar copy paths
        layers = [torch.nn.Linear(dims[i], dims[i + 1], bias=(i == 1)) for i in range(4)]
        m = torch.nn.Sequential(*layers).to(dev)
        mpp = MixedPrecisionPolicy(param_dtype=BF, reduce_dtype=torch.float32)
        for layer in m:
            fully_shard(layer, mp_policy=mpp, reshard_after_forward=reshard)
        fully_shard(m, mp_policy=mpp)
        if factor is not None:
            for layer in m:
                layer.set_gradient_divide_factor(factor)
        return m

    g = torch.Generator(device="cpu").manual_seed(40 + rank)
    x = torch.randn(32, 256, generator=g).to(dev)

    def grads(m, steps=2):
        for _ in range(steps):  # the second step exercises buffer reuse
            for p in m.parameters():
                p.grad = None
            m(x).float().square().mean().backward()
        torch.cuda.synchronize()
        return {n: p.grad.to_local().clone() for n, p in m.named_parameters()}

    res = {}
    for factor in (None, 4.0):
        ref = grads(build(factor=factor))
        got = {}
        for mode, fuse, reshard in (("push", True, True), ("push", True, False), ("pull", False, True), ("f32", True, True)):
            m = build(reshard=reshard, factor=factor)
            install_fsdp_comm(m, group=group, rs_mode=mode, fuse_copy_out=fuse)
            got[(mode, fuse, reshard)] = grads(m)
            for n, gr in got[(mode, fuse, reshard)].items():
                # same bf16 GEMMs; only the fp32 summation order of the reduce-scatter differs from NCCL's ring
                if not torch.allclose(gr, ref[n], atol=1e-6, rtol=1e-4):
                    _fail(f"FSDP2 gradient of {n} (rs_mode={mode}, fused copy-out={fuse}, reshard={reshard}, divide factor={factor}): "
                          f"max abs diff {(gr - ref[n]).abs().max().item():.3e}")
        keys = list(got)
        for k in keys[1:]:
            for n in got[k]:
                if not torch.equal(got[k][n], got[keys[0]][n]):
                    _fail(f"FSDP2 gradient of {n} differs between comm modes {keys[0]} and {k}")
    res["fsdp2_grads_vs_nccl"] = "rtol 1e-4 (summation order only)"
    res["fsdp2_grads_across_modes"] = "bit-exact"
    res["fsdp2_gradient_divide_factor"] = "rtol 1e-4"
    return res


def run_all(symm: SymmetricMemory, dev, fsdp: bool = True, ep: bool = True) -> dict:
    out = {"allgather": check_allgather(symm, dev), "allgather_fused_copy_out": check_allgather_scatter(symm, dev)}
    out.update(check_reduce_scatter(symm, dev))
    out["reduce_scatter_fused_copy_in"] = check_reduce_scatter_push(symm, dev)
    out["ulysses_all_to_all"] = check_ulysses(symm, dev)
    out["ulysses_all_to_all_images"] = check_images(symm, dev)
    if ep:
        out["ep_dispatch_combine"] = check_ep_dispatch(symm, dev)
    if fsdp:
        out.update(check_fsdp(dev, symm.world, symm.group))
    return out
