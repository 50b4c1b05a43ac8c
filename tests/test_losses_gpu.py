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


# ---- the detection losses built on it against the UNMODIFIED reference file ------------------

def _loss_inputs(seed, b=4, n=20000, s=1024, k=256, k2=64, nh=1, ns=18, nc=18):
    """A data_dict with the keys lib/loss_helper.py:25-193 reads, shaped like the detector's outputs and the
    ScanNet loader's labels (lib/dataset.py), random but with real structure: some seeds on objects, some
    proposals near GT centres, padded GT boxes masked out."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *sh: torch.randn(*sh, generator=g)
    gt_center = torch.rand(b, k2, 3, generator=g) * 6 - 3
    nbox = torch.randint(3, k2 // 2, (b,), generator=g)
    box_mask = (torch.arange(k2)[None] < nbox[:, None]).float()
    gt_center = gt_center * box_mask[..., None]                        # padded boxes sit at the origin
    agg = gt_center[:, torch.randint(0, k2, (k,), generator=g)] + r(b, k, 3) * 0.35   # near / grey zone / far mix
    d = {
        "seed_xyz": r(b, s, 3), "seed_inds": torch.randint(0, n, (b, s), generator=g).int(),
        "vote_xyz": r(b, s, 3), "vote_label": r(b, n, 9) * 0.5,
        "vote_label_mask": (torch.rand(b, n, generator=g) < 0.4).long(),
        "aggregated_vote_xyz": agg, "center": agg + r(b, k, 3) * 0.05,
        "center_label": gt_center, "box_label_mask": box_mask,
        "objectness_scores": r(b, k, 2), "heading_scores": r(b, k, nh), "heading_residuals_normalized": r(b, k, nh),
        "size_scores": r(b, k, ns), "size_residuals_normalized": r(b, k, ns, 3), "sem_cls_scores": r(b, k, nc),
        "heading_class_label": torch.randint(0, nh, (b, k2), generator=g),
        "heading_residual_label": r(b, k2) * 0.1,
        "size_class_label": torch.randint(0, ns, (b, k2), generator=g),
        "size_residual_label": r(b, k2, 3) * 0.1,
        "sem_cls_label": torch.randint(0, nc, (b, k2), generator=g),
    }
    return {key: v.cuda() for key, v in d.items()}


class _Config(object):
    num_heading_bin, num_size_cluster, num_class = 1, 18, 18
    mean_size_arr = torch.rand(18, 3, generator=torch.Generator().manual_seed(5)).numpy() + 0.3


@pytest.mark.parametrize("seed", [0, 1])
def test_detection_losses_match_the_unmodified_reference_loss_helper(seed):
    """bridgeqa_b200.loss_helper vs lib/loss_helper.py:25-193 run as is (oracle/_ref/ref_tree) on the same GPU
    and inputs: labels / masks / assignments identical, every loss term to 1e-6, and the gradients reaching
    vote_xyz, center and every score tensor to 1e-5."""
    from oracle import ref_ext
    ref = ref_ext.load_reference_loss_helper()
    if ref is None:
        pytest.skip("oracle/_ref/ref_tree (staged reference loss_helper.py) not built")
    from bridgeqa_b200 import loss_helper as mine
    grads_of = ["vote_xyz", "center", "objectness_scores", "heading_scores", "heading_residuals_normalized",
                "size_scores", "size_residuals_normalized", "sem_cls_scores"]

    def run(mod, cfg):
        d = _loss_inputs(seed)
        for key in grads_of:
            d[key].requires_grad_(True)
        vote = mod.compute_vote_loss(d)
        obj, label, mask, assign = mod.compute_objectness_loss(d)
        d["objectness_label"], d["objectness_mask"], d["object_assignment"] = label, mask, assign
        terms = mod.compute_box_and_sem_cls_loss(d, cfg)
        total = vote + obj + sum(terms)
        total.backward()
        return [vote, obj] + list(terms), (label, mask, assign), [d[key].grad for key in grads_of]
    t_ref, l_ref, g_ref = run(ref, _Config())
    t_mine, l_mine, g_mine = run(mine, _Config())
    assert int(l_ref[0].sum()) > 0 and float(l_ref[1].sum()) < l_ref[1].numel()      # positives and a grey zone exist
    for a, b_ in zip(l_mine, l_ref):
        assert torch.equal(a, b_.to(a.dtype))
    for a, b_ in zip(t_mine, t_ref):
        torch.testing.assert_close(a, b_, rtol=1e-6, atol=1e-7)
    for name, a, b_ in zip(grads_of, g_mine, g_ref):
        torch.testing.assert_close(a, b_, rtol=1e-5, atol=1e-7, msg=name)
    # the composition (lib/loss_helper.py:354-464, detection terms, default weights)
    d = _loss_inputs(seed)
    loss, d = mine.get_detection_loss(d, _Config())
    want = 10.0 * (t_ref[0] + t_ref[1] + (t_ref[2] + 0.1 * t_ref[3] + t_ref[4] + 0.1 * t_ref[5] + t_ref[6]) + t_ref[7])
    torch.testing.assert_close(loss, want.detach(), rtol=1e-6, atol=1e-6)
