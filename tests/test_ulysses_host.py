"""Host-side geometry of the Ulysses exchange: the descriptors the kernel consumes, interpreted in numpy,
reproduce the reference-pinned oracle (tests/golden/multirank.pt) bit for bit."""
import numpy as np
import pytest
import torch

from oracle import comm as o_comm
from veomni_b200._lib import VB200Error
from veomni_b200.ulysses import a2a_plan, images_chunks


def _emulate(xs, scatter_dim, gather_dim):
    """Run the kernel's addressing (p2p.cu all_to_all_kernel) on host byte buffers."""
    P = len(xs)
    item = xs[0].element_size()
    outs = []
    for r in range(P):
        out_shape, (srs, srow, dps, drow, rows, seg) = a2a_plan(tuple(xs[r].shape), scatter_dim, gather_dim, P, item)
        nbytes = int(np.prod(out_shape)) * item
        dst = np.zeros(nbytes, dtype=np.uint8)
        for p in range(P):
            src = xs[p].contiguous().view(torch.uint8).numpy().reshape(-1)
            for row in range(rows):
                s = r * srs + row * srow
                d = p * dps + row * drow
                dst[d : d + seg] = src[s : s + seg]
        outs.append(torch.from_numpy(dst).view(xs[0].dtype).view(out_shape))
    return outs


def test_descriptors_reproduce_reference(golden):
    ranks = golden("multirank.pt")["ranks"]
    xs = [r["ulysses"]["x"] for r in ranks]
    got = _emulate(xs, scatter_dim=1, gather_dim=0)
    for r, rk in enumerate(ranks):
        assert torch.equal(got[r], rk["ulysses"]["gathered"])
    back = _emulate(got, scatter_dim=0, gather_dim=1)
    for r in range(len(ranks)):
        assert torch.equal(back[r], xs[r])


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("shape,sd,gd", [((6, 16, 8), 1, 0), ((1, 6, 16, 8), 2, 1), ((16, 8, 8), 0, 1), ((1, 32, 8, 16), 1, 2)])
def test_descriptors_vs_oracle(P, shape, sd, gd):
    g = torch.Generator().manual_seed(P)
    xs = [torch.randn(*shape, generator=g).to(torch.bfloat16) for _ in range(P)]
    ref = o_comm.all_to_all_tensor(xs, sd, gd)
    got = _emulate(xs, sd, gd)
    for r in range(P):
        assert torch.equal(got[r], ref[r])


def test_unsupported_layouts_fail_loudly():
    with pytest.raises(VB200Error):
        a2a_plan((4, 6, 8, 16), 3, 1, 2, 2)  # non-adjacent dims
    with pytest.raises(VB200Error):
        a2a_plan((2, 6, 8, 16), 2, 1, 2, 2)  # batch > 1
    with pytest.raises(VB200Error):
        a2a_plan((6, 7, 16), 1, 0, 2, 2)  # heads not divisible


@pytest.mark.parametrize("world,seed", [(2, 0), (4, 1), (8, 2)])
def test_image_row_exchange_chunks_vs_oracle(world, seed):
    """The block list all_to_all_images hands to vb200_chunk_pull, interpreted on host byte buffers, reproduces the
    definition of the reference's uneven row exchange (oracle.comm.all_to_all_rows) — zero-row blocks included — and the
    transposed split matrix gives the backward exchange."""
    g = torch.Generator().manual_seed(seed)
    splits = torch.randint(0, 5, (world, world), generator=g)
    splits[0, world - 1] = 0
    H = 24  # 48-byte rows
    xs = [torch.randn(int(splits[s].sum()), H, generator=g).to(torch.bfloat16) for s in range(world)]
    ref = o_comm.all_to_all_rows(xs, splits.tolist())
    row = H * 2

    def pull(bufs, mat, r):
        ch = images_chunks(mat, r, row)
        assert ch.dtype == torch.int64 and ch.shape == (world, 4) and ch.is_contiguous()
        out = np.zeros(int(mat[:, r].sum()) * row, dtype=np.uint8)
        for src_off, dst_off, nbytes, peer in ch.tolist():
            assert src_off % 16 == 0 and dst_off % 16 == 0 and nbytes % 16 == 0
            src = bufs[peer].contiguous().view(torch.uint8).numpy().reshape(-1)
            out[dst_off : dst_off + nbytes] = src[src_off : src_off + nbytes]
        if out.size == 0:
            return torch.zeros(0, H, dtype=torch.bfloat16)
        return torch.from_numpy(out).view(torch.bfloat16).view(-1, H)

    outs = [pull(xs, splits, r) for r in range(world)]
    for r in range(world):
        assert torch.equal(outs[r], ref[r])
    back = [pull(outs, splits.t().contiguous(), r) for r in range(world)]  # backward == the exchange with the roles swapped
    for r in range(world):
        assert torch.equal(back[r], xs[r])
