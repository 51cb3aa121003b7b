"""Host-side parallel logic on CPU: mesh layout (world_size-2 gloo), EP rank matrix, ParallelPlan slicing,
symmetric-memory allocator bookkeeping."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from veomni_b200.parallel_plan import ParallelPlan, check_fqn_match, qwen3_moe_parallel_plan
from veomni_b200.parallel_state import init_para_mesh_matrix


def test_ep_rank_matrix_matches_reference_layout():
    # reference init_para_mesh_matrix (parallel_state.py:50-73): EP ranks consecutive unless `outside`
    assert init_para_mesh_matrix(2, 4).tolist() == [[0, 2, 4, 6], [1, 3, 5, 7]]
    assert init_para_mesh_matrix(2, 4, para_outside=True).tolist() == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert init_para_mesh_matrix(8, 1).tolist() == [[r] for r in range(8)]


def test_fqn_patterns():
    assert check_fqn_match("model.layers.*.mlp.experts.gate_up_proj", "model.layers.17.mlp.experts.gate_up_proj")
    assert not check_fqn_match("model.layers.*.mlp.experts.gate_up_proj", "model.layers.1.2.mlp.experts.gate_up_proj")
    assert not check_fqn_match("model.layers.*.mlp.experts.down_proj", "model.layers.0.mlp.experts.gate_up_proj")


def test_parallel_plan_slices_expert_weights():
    class Experts(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_up_proj = torch.nn.Parameter(torch.arange(8 * 4 * 2, dtype=torch.float32).view(8, 4, 2))
            self.down_proj = torch.nn.Parameter(torch.zeros(8, 2, 2))

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Module()
            layer = torch.nn.Module()
            layer.mlp = torch.nn.Module()
            layer.mlp.experts = Experts()
            layer.norm = torch.nn.LayerNorm(2)
            self.model.layers = torch.nn.ModuleList([layer])

    m = M()
    full = m.model.layers[0].mlp.experts.gate_up_proj.data.clone()
    info = qwen3_moe_parallel_plan().apply(m, ep_size=4, ep_rank=2)
    got = m.model.layers[0].mlp.experts.gate_up_proj
    assert got.shape == (2, 4, 2) and torch.equal(got.data, full[4:6])
    assert info["model.layers.0.mlp.experts.gate_up_proj"].placement.dim == 0
    assert not hasattr(info["model.layers.0.norm.weight"].placement, "dim")  # Replicate
    assert all(hasattr(p, "spec_info") for p in m.parameters())
    with pytest.raises(AssertionError):
        ParallelPlan({"ep": {"w": __import__("torch").distributed._tensor.Shard(0)}}).shard_tensor(torch.zeros(3, 2), "w", 2, 0)


def _mesh_worker(rank, world, path, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    store = dist.FileStore(path, world)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    from veomni_b200.parallel_state import init_parallel_state

    ps = init_parallel_state(dp_size=1, ulysses_size=2, ep_size=2, device_type="cpu")
    res = {
        "fsdp_ranks": dist.get_process_group_ranks(ps.fsdp_group),
        "ulysses_ranks": dist.get_process_group_ranks(ps.ulysses_group),
        "ep_ranks": dist.get_process_group_ranks(ps.ep_group),
        "sp": ps.sp_enabled, "ep": ps.ep_enabled, "div": ps.extra_parallel_gradient_divide_factor("ep"),
    }
    torch.save(res, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_parallel_state_mesh_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_mesh_worker, args=(2, os.path.join(d, "store"), d), nprocs=2, join=True)
        for r in range(2):
            res = torch.load(os.path.join(d, f"r{r}.pt"))
            # Ulysses ranks are also FSDP shard ranks (parallel_state.py:87,95-96)
            assert res["fsdp_ranks"] == [0, 1] and res["ulysses_ranks"] == [0, 1] and res["ep_ranks"] == [0, 1]
            assert res["sp"] and res["ep"] and res["div"] == 2


def test_symmetric_allocator_bookkeeping():
    """First-fit free list with coalescing (pure host logic of veomni_b200.symm.Arena)."""
    from veomni_b200._lib import VB200Error
    from veomni_b200.symm import Arena

    ar = Arena(1024, 4096)
    a, sa = ar.alloc(1000)  # rounded up to 256-byte granules
    b, sb = ar.alloc(1024)
    c, sc = ar.alloc(2048)
    assert (a, b, c) == (1024, 2048, 3072) and (sa, sb, sc) == (1024, 1024, 2048) and ar.free == []
    with pytest.raises(VB200Error):
        ar.alloc(1)
    ar.release(b, sb)
    ar.release(a, sa)
    assert ar.free == [(1024, 2048)]  # coalesced
    assert ar.alloc(2048)[0] == 1024
    ar.release(1024, 2048)
    ar.release(c, sc)
    assert ar.free == [(1024, 4096)] and ar.live == 0


def test_pack_plan_reproduces_chunk_cat_layout():
    """The geometry the bf16 pack kernel uses == torch._chunk_cat(dim=0, num_chunks=world) as FSDP2's
    foreach_reduce_scatter_copy_in calls it (ragged dim-0 sizes are zero-padded per parameter)."""
    import torch

    from veomni_b200.fsdp_comm import pack_plan

    g = torch.Generator().manual_seed(0)
    for world in (2, 4, 8):
        shapes = [(10, 4), (3,), (7, 2, 2), (16, 8), (1, 5)]
        grads = [torch.randn(*s, generator=g) for s in shapes]
        plan, row = pack_plan(shapes, world)
        ref = torch.empty(world, row)
        torch._chunk_cat(grads, dim=0, num_chunks=world, out=ref)
        out = torch.full((world, row), float("nan"))
        for t, (numel, chunk, off) in zip(grads, plan):
            flat = torch.cat([t.reshape(-1), torch.zeros(chunk * world - numel)])  # el -> (rank, within) as in the kernel
            for el in range(chunk * world):
                r = el // chunk
                out[r, off + el - r * chunk] = flat[el]
        assert torch.equal(out, ref), world


def _fsdp_patch_worker(rank, world, path, outdir):
    """FSDP2 on gloo/CPU with and without the copy-in patches installed: CPU tensors must take PyTorch's own path."""
    import torch
    import torch.distributed as dist
    from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

    store = dist.FileStore(path, world)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)

    def grads(patched: bool):
        torch.manual_seed(5)
        m = torch.nn.Sequential(torch.nn.Linear(24, 36, bias=False), torch.nn.Linear(36, 10))
        mp_policy = MixedPrecisionPolicy(param_dtype=torch.bfloat16, reduce_dtype=torch.float32)
        for layer in m:
            fully_shard(layer, mp_policy=mp_policy)
        fully_shard(m, mp_policy=mp_policy)
        if patched:
            from veomni_b200 import fsdp_comm

            fsdp_comm._patch_copy_in()
        x = torch.randn(8, 24, generator=torch.Generator().manual_seed(9 + rank))
        m(x).float().square().mean().backward()
        return [p.grad.to_local().clone() for p in m.parameters()]

    a, b = grads(False), grads(True)
    ok = all(torch.equal(x, y) for x, y in zip(a, b))
    torch.save(ok, os.path.join(outdir, f"ok{rank}.pt"))
    dist.destroy_process_group()


def test_copy_in_patches_leave_the_cpu_path_alone_world2_gloo():
    import torch
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_fsdp_patch_worker, args=(2, os.path.join(d, "store"), d), nprocs=2, join=True)
        assert all(torch.load(os.path.join(d, f"ok{r}.pt")) for r in range(2))


def _sp_loss_worker(rank, world, path, outdir):
    import torch
    import torch.distributed as dist

    from veomni_b200.cross_entropy import _ReduceLoss

    store = dist.FileStore(path, world)
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)
    f = torch.load(os.path.join(os.path.dirname(__file__), "golden", "multirank.pt"), weights_only=False)["ranks"][rank]["sp_loss"]
    ok = True
    for case in f:
        loss_in = torch.tensor(case["loss"], requires_grad=True)
        out = _ReduceLoss.apply(loss_in * 1.0, torch.tensor(case["n_valid"]), dist.group.WORLD)
        (g,) = torch.autograd.grad(out * 3.0, loss_in)
        ok = ok and torch.allclose(out.detach(), case["reduced"], atol=1e-6) and torch.allclose(g, case["grad"], atol=1e-6)
    torch.save(ok, os.path.join(outdir, f"ok{rank}.pt"))
    dist.destroy_process_group()


def test_sp_loss_reduce_matches_reference_world2_gloo():
    """The SP loss reduction of the causal-LM loss wrapper against the reference's ReduceLoss outputs (fixture made by the
    reference on a 2-rank gloo group): token-weighted mean, empty ranks, all-empty group, and the backward factor."""
    import torch
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_sp_loss_worker, args=(2, os.path.join(d, "store"), d), nprocs=2, join=True)
        assert all(torch.load(os.path.join(d, f"ok{r}.pt")) for r in range(2))


def test_ep_plan_chunks_matches_the_loop_restatement():
    """veomni_b200.ep.plan_chunks (vectorised) == the per-(expert, source) loops that define the EP exchange geometry
    (moe_layer.py:30-69 preprocess + moe_utils.py:81-99 sort_chunks_by_idxs), for every rank, with empty blocks."""
    import torch

    from veomni_b200.ep import plan_chunks

    g = torch.Generator().manual_seed(0)
    for ep, el in ((2, 4), (4, 3), (8, 16)):
        E = ep * el
        counts = torch.randint(0, 7, (ep, E), generator=g, dtype=torch.int64)
        counts[0, 1] = 0
        counts[:, E - 1] = 0  # an expert nobody picked
        row = 64
        excl = torch.cumsum(counts, dim=1) - counts
        for r in range(ep):
            ins, outs, total, cumsum, fwd, bwd = plan_chunks(counts, r, row)
            assert ins == counts[r].view(ep, el).sum(1).tolist()
            assert outs == counts[:, r * el:(r + 1) * el].sum(1).tolist()
            exp_fwd, dst = [], 0
            for le in range(el):
                e = r * el + le
                for s in range(ep):
                    n = int(counts[s, e])
                    exp_fwd.append([int(excl[s, e]) * row, dst * row, n * row, s])
                    dst += n
            assert total == dst
            assert fwd.tolist() == exp_fwd
            assert cumsum.tolist() == torch.cumsum(counts[:, r * el:(r + 1) * el].sum(0), 0).tolist()
            exp_bwd = []
            tot = counts.sum(0)
            for p in range(ep):
                base = 0
                for le in range(el):
                    e = p * el + le
                    exp_bwd.append([(base + int(counts[:r, e].sum())) * row, int(excl[r, e]) * row, int(counts[r, e]) * row, p])
                    base += int(tot[e])
            assert bwd.tolist() == exp_bwd


def _wrap_worker(rank, world, ep, rep, path, outdir):
    """Our build_parallelize_model on the host toy Qwen3-MoE, 2 or 4 gloo ranks, NCCL-free (b200_comm off: CPU)."""
    import json
    import sys

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    torch.set_num_threads(1)
    dist.init_process_group("gloo", store=dist.FileStore(path, world), rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_wrap_structure import TOY, describe
    from veomni_b200.host_qwen3_moe import Qwen3MoeConfig, Qwen3MoeForCausalLM
    from veomni_b200.parallel_state import init_parallel_state
    from veomni_b200.parallelize import build_parallelize_model

    init_parallel_state(dp_size=world, dp_replicate_size=rep, dp_shard_size=world // rep, ulysses_size=1, ep_size=ep,
                        device_type="cpu")
    cfg = Qwen3MoeConfig(**{k: TOY[k] for k in (
        "vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
        "head_dim", "rms_norm_eps", "rope_theta", "tie_word_embeddings", "initializer_range", "num_experts",
        "num_experts_per_tok", "moe_intermediate_size", "norm_topk_prob")})
    with torch.device("meta"):
        model = Qwen3MoeForCausalLM(cfg)
    model = build_parallelize_model(model, init_device="meta", b200_comm=False, enable_reshard_after_forward=True,
                                    enable_gradient_checkpointing=True, enable_forward_prefetch=True)
    res = describe(model)
    res["finite"] = all(bool(torch.isfinite(p.to_local()).all()) for p in model.parameters())
    with open(os.path.join(outdir, f"r{rank}.json"), "w") as fh:
        json.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ep,rep", [(2, 2, 1), (4, 2, 1), (2, 1, 1), (4, 1, 2)])
def test_wrap_structure_matches_the_reference_gloo(world, ep, rep):
    """``build_parallelize_model`` (EP slice + experts Shard(1) on ep_fsdp + bottom-up FSDP2 + prefetch lists + meta init)
    produces the same FSDP2 structure — EP with ep_fsdp 1 and 2, dense, and HSDP (2 replicas x 2 shards) — as the reference's
    own function on the same toy Qwen3-MoE
    (tests/golden/wrap_structure.json, generated by running the unmodified reference on gloo: make_wrap_structure.py):
    same units on the same meshes with the same placements / local shapes / dtypes / divide factors / prefetch targets."""
    import json

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wrap_structure.json")))
    run = next(r for r in ref["runs"] if (r["world"], r["ep_size"], r["dp_replicate"]) == (world, ep, rep))
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_wrap_worker, args=(world, ep, rep, os.path.join(d, "store"), d), nprocs=world, join=True)
        for r in range(world):
            mine, want = json.load(open(os.path.join(d, f"r{r}.json"))), run["ranks"][r]
            assert mine["finite"] and mine["all_dtensor"] and mine["has_clip"] and mine["grad_ckpt"]
            assert mine["param_dtypes"] == want["param_dtypes"] == ["torch.float32"]
            assert mine["ep_sliced_fqns"] == want["ep_sliced_fqns"]
            assert [u["module"] for u in mine["units"]] == [u["module"] for u in want["units"]]
            for a, b in zip(mine["units"], want["units"]):
                for k in ("cls", "reshard_after_forward", "auto_reshard_after_forward", "gradient_divide_factor", "param_dtype",
                          "reduce_dtype", "backward_prefetch"):
                    assert a[k] == b[k], (a["module"], k, a[k], b[k])
                assert [s for _n, s in a["mesh"]] == [s for _n, s in b["mesh"]], (a["module"], a["mesh"], b["mesh"])
                assert sorted(map(json.dumps, a["params"])) == sorted(map(json.dumps, b["params"])), a["module"]
                if ep > 1:
                    assert a["forward_prefetch"] == b["forward_prefetch"], (a["module"], a["forward_prefetch"])
            if ep == 1:
                # the reference sets no explicit lists without EP (FSDP2's implicit prefetch); ours chains root -> layer 0 ->
                # layer 1 ... so that the overlap does not depend on host timing (parallelize.py, "explicit forward prefetch")
                chain = {u["module"]: u["forward_prefetch"] for u in mine["units"]}
                L = ref["toy_config"]["num_hidden_layers"]
                assert chain[""] == ["model.layers.0"] and chain[f"model.layers.{L - 1}"] == []
                assert all(chain[f"model.layers.{i}"] == [f"model.layers.{i + 1}"] for i in range(L - 1))
