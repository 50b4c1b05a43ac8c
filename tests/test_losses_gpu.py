"""`nn_distance` (utils/nn_distance.py:25-52) on the sm_100a kernel vs the
reference's torch expression, restated below line by line (it is pure torch): forward bit-exact,
gradients to both point sets equal."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from bridgeqa_b200 import nn_distance as nd  # noqa: E402


def reference_nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    # utils/nn_distance.py:38-52
    n, m = pc1.shape[1], pc2.shape[1]
    pc_diff = pc1.unsqueeze(2).repeat(1, 1, m, 1) - pc2.unsqueeze(1).repeat(1, n, 1, 1)
    if l1smooth:
        pc_dist = torch.sum(nd.huber_loss(pc_diff, delta), dim=-1)
    elif l1:
        pc_dist = torch.sum(torch.abs(pc_diff), dim=-1)
    else:
        pc_dist = torch.sum(pc_diff ** 2, dim=-1)
    dist1, idx1 = torch.min(pc_dist, dim=2)
    dist2, idx2 = torch.min(pc_dist, dim=1)
    return dist1, idx1, dist2, idx2


CASES = [
    # (B, N, M, kwargs) -- the three call sites of lib/loss_helper.py (vote loss: B*num_seed tiny sets, L1;
    # objectness / box loss: 256 proposals vs 128 padded GT centres) and ragged sizes around the tile
    (16 * 1024, 1, 3, dict(l1=True)), (16, 256, 128, {}), (3, 257, 511, dict(l1smooth=True, delta=0.15)),
    (2, 1, 1, {}), (4, 700, 5, dict(l1=True)), (2, 64, 300, dict(l1smooth=True, delta=1.0)),
]


@pytest.mark.parametrize("b,n,m,kw", CASES)
def test_nn_distance_matches_reference_expression(b, n, m, kw):
    g = torch.Generator(device="cpu").manual_seed(b * 1000 + n + m)
    pc1 = torch.randn(b, n, 3, generator=g).cuda()
    pc2 = torch.randn(b, m, 3, generator=g).cuda()
    if m > 4:
        pc2[:, 3] = pc2[:, 1]                       # duplicates: torch.min must keep the first
    if n > 4:
        pc1[:, 2] = pc1[:, 0]
    pc2[:, -1] = 0.0                                # like the zero-padded GT centres
    a1, a2 = pc1.clone().requires_grad_(True), pc2.clone().requires_grad_(True)
    r1, r2 = pc1.clone().requires_grad_(True), pc2.clone().requires_grad_(True)
    got = nd.nn_distance(a1, a2, **kw)
    want = reference_nn_distance(r1, r2, **kw)
    for x, y in zip(got, want):
        assert x.dtype == y.dtype and torch.equal(x, y)
    w1 = torch.randn_like(got[0])
    w2 = torch.randn_like(got[2])
    ((got[0] * w1).sum() + (got[2] * w2).sum()).backward()
    ((want[0] * w1).sum() + (want[2] * w2).sum()).backward()
    # gradients are sums of up to n fp32 terms per point, accumulated in a different (atomic,
    # nondeterministic) order than torch's scatter: the 1e-4 bar of every scatter-add on this path
    torch.testing.assert_close(a1.grad, r1.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(a2.grad, r2.grad, rtol=1e-4, atol=1e-5)


def test_nn_distance_errors():
    with pytest.raises(RuntimeError):
        nd.nn_distance(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3).cuda())      # CPU tensor
    with pytest.raises(RuntimeError):
        nd.nn_distance(torch.zeros(1, 4, 2).cuda(), torch.zeros(1, 4, 2).cuda())   # not xyz
