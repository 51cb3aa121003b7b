"""Host-side geometry of the Ulysses exchange: the descriptors the kernel consumes, interpreted in numpy,
reproduce the reference-pinned oracle (tests/golden/multirank.pt) bit for bit."""
import numpy as np
import pytest
import torch

from oracle import comm as o_comm
from veomni_b200._lib import VB200Error
from veomni_b200.ulysses import a2a_plan


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
