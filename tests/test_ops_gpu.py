"""GPU parity of the nine operators: sm_100a kernels (through the C ABI) vs the C oracle
(bit-exact for indices, gathers and the fp32 interpolation), vs the reference's own CUDA
extension when oracle/_ref is present, plus size-independent properties at the full
BASELINE sizes.  Mirrors how the reference's only test drives the ops
(lib/pointnet2/pointnet2_test.py:18-30) -- through pointnet2_utils.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from bridgeqa_b200 import ext, pointnet2_utils as pu, synthetic  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scenes(b, n, first=0):
    return synthetic.make_batch(b, n, 0, first_scene=first)[..., :3].contiguous().numpy()


# ------------------------------------------------------------------------ FPS ---

FPS_CASES = [
    # (B, N, npoint)  -- covers cluster sizes 1/2/4/8/16, several P, n < 512, ragged n
    (2, 1, 1), (2, 7, 5), (3, 100, 33), (2, 511, 64), (2, 512, 256), (4, 1024, 512),
    (2, 2048, 1024), (2, 3000, 300), (1, 8192, 400), (2, 8193, 200), (2, 12345, 300),
    (2, 20000, 512), (2, 40000, 256), (1, 50000, 128), (1, 70000, 96), (1, 100000, 64),
]


@pytest.mark.parametrize("b,n,m", FPS_CASES)
def test_fps_matches_oracle(oracle_ops, b, n, m):
    xyz = scenes(b, n, first=b + n % 17)
    want = oracle_ops.furthest_point_sampling(xyz, m)
    got, new_xyz = ext.furthest_point_sampling(dev(xyz), m, return_xyz=True)
    got = got.cpu().numpy()
    assert got.dtype == np.int32 and got.shape == (b, m)
    np.testing.assert_array_equal(got, want)
    # fused gather epilogue == xyz[inds]
    np.testing.assert_array_equal(new_xyz.cpu().numpy(),
                                  np.take_along_axis(xyz, want[..., None].astype(np.int64), 1))


def test_fps_ties_and_skips(oracle_ops):
    """Heavy exact duplication (ties everywhere) + many |p|^2 <= 1e-3 points + index-0 skip."""
    rng = np.random.default_rng(5)
    base = rng.uniform(-2, 2, (2, 300, 3)).astype(np.float32)
    xyz = base[:, rng.integers(0, 300, 6000)]                 # every point ~20 copies
    xyz[:, ::7] = rng.uniform(-0.01, 0.01, xyz[:, ::7].shape).astype(np.float32)
    xyz[0, 0] = 0.0                                           # the seed point itself is skippable
    want = oracle_ops.furthest_point_sampling(xyz, 512)
    got = ext.furthest_point_sampling(dev(xyz), 512).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_fps_grid_ties(oracle_ops):
    """Integer lattice: huge sets of exactly equal distances -> pure tie-break test."""
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(8), indexing="ij"), -1)
    xyz = g.reshape(1, -1, 3).astype(np.float32) + 1.0
    rng = np.random.default_rng(1)
    xyz = np.stack([xyz[0], xyz[0][rng.permutation(xyz.shape[1])]], 0)
    want = oracle_ops.furthest_point_sampling(xyz, 700)
    got = ext.furthest_point_sampling(dev(xyz), 700).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_fps_all_points_skipped(oracle_ops):
    xyz = np.full((2, 1000, 3), 0.001, dtype=np.float32)
    got = ext.furthest_point_sampling(dev(xyz), 16).cpu().numpy()
    np.testing.assert_array_equal(got, np.zeros((2, 16), np.int32))
    np.testing.assert_array_equal(got, oracle_ops.furthest_point_sampling(xyz, 16))


def test_fps_full_size_properties():
    """B=16, N=40000 -> 2048 (BASELINE size): indices in range, distinct (scene points are
    distinct apart from the 1 % duplicates, which FPS never re-picks while unpicked points
    remain), first index 0, and the min-distance sequence is non-increasing."""
    xyz = scenes(16, 40000)
    inds, new_xyz = ext.furthest_point_sampling(dev(xyz), 2048, return_xyz=True)
    inds = inds.cpu().numpy()
    assert inds.min() >= 0 and inds.max() < 40000
    assert (inds[:, 0] == 0).all()
    p = new_xyz.cpu().numpy().astype(np.float64)
    for s in range(0, 16, 5):
        assert len(np.unique(inds[s])) == 2048
        # distance of sample j to the set of earlier samples must not increase with j
        d = np.full(2048, np.inf)
        last = []
        for j in range(1, 2048, 64):
            dj = np.min(np.sum((p[s, :j] - p[s, j]) ** 2, -1))
            last.append(dj)
        assert all(a >= b - 1e-6 for a, b in zip(last, last[1:]))


def test_fps_vs_reference_extension(ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    for (b, n, m) in [(4, 40000, 2048), (4, 2048, 1024), (3, 1024, 512), (2, 512, 256), (2, 300, 64)]:
        xyz = dev(scenes(b, n, first=3))
        want = ref_ext.furthest_point_sampling(xyz, m)
        got = ext.furthest_point_sampling(xyz, m)
        assert torch.equal(got, want), (b, n, m)


def _fps_grid(xyz, m, radius=0.2, lean=False):
    from bridgeqa_b200 import fused
    grid = fused.prebuild_ball_query_grid(xyz, radius, inline=True)
    with fused.lean_sampling(lean):        # lean: the throughput variant (coordinates in shared memory)
        return fused.furthest_point_sample_grid(xyz, m, grid)


FPS_GRID_CASES = [
    # (B, N, npoint, build radius) -- cluster sizes 1..16, ragged n, tiny and huge cells
    (2, 512, 256, 0.4), (3, 1000, 333, 0.2), (2, 2048, 1024, 0.4), (2, 5000, 600, 0.2),
    (2, 12345, 300, 0.05), (2, 20000, 512, 0.2), (4, 40000, 700, 0.2), (1, 40000, 2048, 3.0),
    (1, 70000, 96, 0.2), (1, 100000, 64, 0.2), (1, 147456, 40, 0.2),
    (1, 60000, 64, 0.2), (1, 90000, 48, 0.2),      # throughput variant: clusters of 5 and 7 CTAs
    # one-SM throughput kernel (fps_stream.cu): 32-point runs up to 32768 points, 64-point runs up to
    # 53440 (its min-distances fill shared memory), the cluster variant beyond
    (1, 32768, 100, 0.2), (1, 32769, 100, 0.2), (1, 53440, 64, 0.2), (1, 53441, 64, 0.2),
]


@pytest.mark.parametrize("lean", [False, True])
@pytest.mark.parametrize("b,n,m,r", FPS_GRID_CASES)
def test_fps_grid_equals_plain_kernel(b, n, m, r, lean):
    xyz = dev(scenes(b, n, first=51))
    want_i, want_x = ext.furthest_point_sampling(xyz, m, return_xyz=True)
    got_i, got_x = _fps_grid(xyz, m, r, lean)
    assert torch.equal(got_i, want_i), (b, n, m)
    assert torch.equal(got_x, want_x)


@pytest.mark.parametrize("lean", [False, True])
def test_fps_grid_matches_oracle_on_ties_and_skips(oracle_ops, lean):
    """Lattice clouds (exact ties every iteration: the carried tie key must reproduce the
    reference's pairwise tree), duplicates, the skip set, non-finite coordinates."""
    rng = np.random.RandomState(7)
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(8), indexing="ij"), -1)
    lattice = g.reshape(-1, 3).astype(np.float32) * np.float32(0.25) + np.float32(0.5)
    pts = np.stack([lattice[rng.permutation(len(lattice))] for _ in range(2)])       # (2, 2048, 3)
    pts[1, 100:140] = pts[1, 100]                     # duplicates
    pts[1, 7] = [0.01, 0.02, 0.01]                    # |p|^2 <= 1e-3: never selectable
    pts[1, 0] = [0.0, 0.0, 0.0]                       # ... but index 0 is always the first sample
    want = oracle_ops.furthest_point_sampling(pts, 700)
    got, _ = _fps_grid(dev(pts), 700, 0.25, lean)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # all points in the skip set -> every sample is index 0
    tiny = (rng.uniform(-0.01, 0.01, size=(1, 600, 3))).astype(np.float32)
    got, got_x = _fps_grid(dev(tiny), 50, 0.2, lean)
    assert (got == 0).all() and torch.equal(got_x[0, 5], dev(tiny)[0, 0])
    # NaN / inf coordinates: same behaviour as the plain kernel
    bad = scenes(1, 3000, first=3)
    bad[0, 11] = [np.nan, 1.0, 1.0]
    bad[0, 12] = [np.inf, 1.0, 1.0]
    want_i = ext.furthest_point_sampling(dev(bad), 200)
    got_i, _ = _fps_grid(dev(bad), 200, 0.2, lean)
    assert torch.equal(got_i, want_i)


def test_fps_prefix_check_and_conditional_sampling(oracle_ops):
    """Re-sampling a cloud that is in sampling order: the parallel proof + conditional chain must
    return exactly what the plain chain returns, for clean scenes (proved: identity prefix),
    scenes with injected ties / duplicates / skip-set points (not provable: chain) and clouds
    that are not in sampling order at all."""
    from bridgeqa_b200 import fused
    b = 6
    xyz = dev(scenes(b, 20000, first=41))
    _, xyz1 = ext.furthest_point_sampling(xyz, 2048, return_xyz=True)
    xyz1 = xyz1.clone()
    xyz1[1, 1500] = xyz1[1, 700]                  # duplicate of a prefix point -> exact tie at step 700
    xyz1[2, 900] = torch.tensor([0.01, 0.01, 0.01], device="cuda")    # reference skip set
    xyz1[3] = xyz1[3, torch.randperm(2048, device="cuda")]            # not in sampling order
    xyz1[4, 1000:1008] = xyz1[4, 1000]            # a block of identical points
    flags = fused.fps_prefix_check(xyz1, 1024)
    assert flags.tolist() == [0, 1, 1, 1, 1, 0], flags.tolist()
    cur = xyz1
    for m in (1024, 512, 256):
        want_i, want_x = ext.furthest_point_sampling(cur, m, return_xyz=True)
        got_i, got_x = fused.furthest_point_sample_cond(cur, m, flags)
        assert torch.equal(got_i, want_i) and torch.equal(got_x, want_x), m
        np.testing.assert_array_equal(got_i.cpu().numpy(),
                                      oracle_ops.furthest_point_sampling(cur.cpu().numpy(), m))
        ident = torch.arange(m, device="cuda", dtype=torch.int32)
        for s in (0, 5):
            assert torch.equal(want_i[s], ident)          # what the proof claims
        cur = got_x


def test_fps_prefix_check_lattice_ties(oracle_ops):
    """Lattice clouds tie constantly; put one in its own sampling order and check again."""
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(8), indexing="ij"), -1)
    pts = (g.reshape(1, -1, 3).astype(np.float32) * np.float32(0.25) + np.float32(0.5))
    order = oracle_ops.furthest_point_sampling(pts, pts.shape[1])[0].astype(np.int64)
    from bridgeqa_b200 import fused
    cloud = dev(pts[:, order])
    flags = fused.fps_prefix_check(cloud, 512)
    got_i, _ = fused.furthest_point_sample_cond(cloud, 512, flags)
    np.testing.assert_array_equal(got_i.cpu().numpy(),
                                  oracle_ops.furthest_point_sampling(pts[:, order], 512))


# ----------------------------------------------------------------- ball query ---

BQ_CASES = [
    # (B, N, M, radius, nsample)
    (2, 1, 1, 0.5, 4), (2, 37, 5, 0.4, 8), (2, 1000, 129, 0.3, 16), (2, 4099, 300, 0.2, 64),
    (2, 2048, 1024, 0.4, 32), (2, 1024, 512, 0.8, 16), (2, 512, 256, 1.2, 16), (1, 1024, 256, 0.3, 16),
    (2, 10000, 700, 0.05, 64), (1, 40000, 2048, 0.2, 64),
]


@pytest.mark.parametrize("b,n,m,r,ns", BQ_CASES)
def test_ball_query_matches_oracle(oracle_ops, b, n, m, r, ns):
    xyz = scenes(b, n, first=11)
    centres = np.take_along_axis(
        xyz, oracle_ops.furthest_point_sampling(xyz, m)[..., None].astype(np.int64), 1)
    want = oracle_ops.ball_query(centres, xyz, r, ns)
    got = ext.ball_query(dev(centres), dev(xyz), r, ns).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_ball_query_empty_balls_and_unaligned(oracle_ops):
    xyz = scenes(3, 1001, first=2)          # 1001*12 B per scene: scenes 1,2 start unaligned
    centres = xyz[:, :50] + np.float32(100.0)   # nothing within reach -> rows of zeros
    centres[:, ::2] = xyz[:, :50:2]
    want = oracle_ops.ball_query(centres, xyz, 0.25, 32)
    got = ext.ball_query(dev(centres), dev(xyz), 0.25, 32).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert (got[:, 1::2] == 0).all()


def test_ball_query_via_utils_arg_order(oracle_ops):
    """pointnet2_utils.ball_query takes (radius, nsample, xyz, new_xyz)."""
    xyz = scenes(2, 3000)
    centres = xyz[:, :128].copy()
    got = pu.ball_query(0.3, 16, dev(xyz), dev(centres)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.ball_query(centres, xyz, 0.3, 16))


def test_ball_query_vs_reference_extension(ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    for (b, n, m, r, ns) in [(4, 40000, 2048, 0.2, 64), (4, 2048, 1024, 0.4, 32), (2, 1024, 256, 0.3, 16)]:
        xyz = dev(scenes(b, n, first=9))
        inds, centres = ext.furthest_point_sampling(xyz, m, return_xyz=True)
        want = ref_ext.ball_query(centres, xyz, r, ns)
        got = ext.ball_query(centres, xyz, r, ns)
        assert torch.equal(got, want), (b, n, m)


def _ball_query_scan(centres, xyz, r, ns):
    """The index-order scan kernel (no workspace => no grid, no segments)."""
    import ctypes
    from bridgeqa_b200 import _native as N
    b, m, _ = centres.shape
    idx = torch.empty((b, m, ns), dtype=torch.int32, device="cuda")
    N.call("bqa_ball_query", b, xyz.size(1), m, ctypes.c_float(r), ns, N.ptr(centres), N.ptr(xyz),
           N.ptr(idx), ctypes.c_void_p(0), N.stream_ptr(xyz.device))
    return idx


GRID_CASES = [
    # (B, N, M, radius, nsample): the cell-grid path (n >= 8192) against the scan kernel
    (2, 8192, 300, 0.2, 64), (3, 20000, 777, 0.35, 16), (2, 40000, 2048, 0.2, 64),
    (1, 40000, 512, 1.5, 128), (1, 100000, 1000, 0.1, 32), (2, 9000, 64, 0.0, 8),
    (1, 9000, 64, 50.0, 16), (2, 12000, 200, 0.2, 1),
]


@pytest.mark.parametrize("b,n,m,r,ns", GRID_CASES)
def test_ball_query_grid_equals_scan(b, n, m, r, ns):
    xyz = dev(scenes(b, n, first=23))
    _, centres = ext.furthest_point_sampling(xyz, m, return_xyz=True)
    centres = centres.clone()
    centres[:, ::7] += 0.013            # off-sample centres too
    assert torch.equal(ext.ball_query(centres, xyz, r, ns), _ball_query_scan(centres, xyz, r, ns))


def test_ball_query_grid_adversarial(oracle_ops):
    """Dense clutter (more hits than the per-warp list holds -> the in-kernel index-order
    fallback), exact duplicates, points on cell borders, far outliers that stretch the bounding
    box, non-finite coordinates, centres outside the box."""
    rng = np.random.RandomState(5)
    n = 10000
    xyz = rng.uniform(-2, 2, size=(2, n, 3)).astype(np.float32)
    xyz[0, 1000:3000] = np.float32([0.5, 0.5, 0.5]) + rng.normal(0, 0.01, size=(2000, 3)).astype(np.float32)
    xyz[0, 4000:4700] = xyz[0, 4000]                           # 700 identical points
    xyz[1, :, :] = np.round(xyz[1] / 0.25) * 0.25              # lattice: everything on cell borders
    xyz[1, 17] = [1e6, -1e6, 3e5]                              # outliers
    xyz[1, 18] = [-4e7, 0, 0]
    xyz[0, 19] = [np.nan, 0.1, 0.1]
    xyz[0, 20] = [np.inf, 0.1, 0.1]
    xyz[0, 21] = [0.1, -np.inf, np.nan]
    centres = np.concatenate([xyz[:, 990:1010], xyz[:, 3995:4005], xyz[:, 10:25],
                              rng.uniform(-3, 3, size=(2, 40, 3)).astype(np.float32)], 1)
    centres[0, -1] = [np.nan, 0, 0]
    centres[1, -1] = [1e6, -1e6, 3e5]
    centres = np.ascontiguousarray(centres)
    for r, ns in [(0.25, 64), (0.05, 16), (0.6, 128), (0.25, 700)]:
        want = oracle_ops.ball_query(centres, xyz, r, ns)
        got = ext.ball_query(dev(centres), dev(xyz), r, ns).cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg=str((r, ns)))
        scan = _ball_query_scan(dev(centres), dev(xyz), r, ns).cpu().numpy()
        np.testing.assert_array_equal(scan, want, err_msg="scan " + str((r, ns)))


def test_ball_query_grid_split_api_any_build_radius():
    """build (on any radius) + search over slices == the one-call form."""
    import ctypes
    from bridgeqa_b200 import _native as N
    b, n, m, ns, r = 2, 30000, 600, 32, 0.3
    xyz = dev(scenes(b, n, first=31))
    _, centres = ext.furthest_point_sampling(xyz, m, return_xyz=True)
    want = _ball_query_scan(centres, xyz, r, ns)
    grid = torch.empty((N.lib().bqa_ball_query_grid_bytes(b, n),), dtype=torch.uint8, device="cuda")
    for build_r in (0.3, 0.02, 5.0):
        N.call("bqa_ball_query_grid_build", b, n, ctypes.c_float(build_r), N.ptr(xyz), N.ptr(grid),
               N.stream_ptr(xyz.device))
        idx = torch.full((b, m, ns), -1, dtype=torch.int32, device="cuda")
        for lo, cnt in [(0, 100), (100, 499), (599, 1)]:
            N.call("bqa_ball_query_grid_search", b, n, m, lo, cnt, ctypes.c_float(r), ns, N.ptr(centres),
                   N.ptr(xyz), N.ptr(idx), N.ptr(grid), N.stream_ptr(xyz.device))
        assert torch.equal(idx, want), build_r


def test_ball_query_full_size_properties():
    """B=16 SA1 size: every returned index is inside the ball (or the row is the all-zero
    empty-ball row), rows are non-decreasing until the back-fill, back-fill == first hit."""
    xyz = dev(scenes(16, 40000))
    inds, centres = ext.furthest_point_sampling(xyz, 2048, return_xyz=True)
    idx = ext.ball_query(centres, xyz, 0.2, 64)
    g = torch.gather(xyz, 1, idx.reshape(16, -1, 1).long().expand(-1, -1, 3)).reshape(16, 2048, 64, 3)
    d2 = ((g - centres[:, :, None]) ** 2).sum(-1)
    assert (d2 < 0.2 * 0.2 + 1e-6).all()          # centre itself is always a hit here
    first = idx[..., :1]
    inc = (idx[..., 1:] > idx[..., :-1]) | (idx[..., 1:] == first)
    assert inc.all()


# ------------------------------------------------------------- gather / group ---

@pytest.mark.parametrize("b,c,n,m", [(2, 3, 40000, 2048), (3, 5, 777, 100), (1, 1, 1, 1), (2, 130, 512, 256)])
def test_gather_and_grad(oracle_ops, b, c, n, m):
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, m)).astype(np.int32)
    got = ext.gather_points(dev(pts), dev(idx)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.gather_points(pts, idx))
    go = rng.standard_normal((b, c, m)).astype(np.float32)
    gg = ext.gather_points_grad(dev(go), dev(idx), n).cpu().numpy()
    np.testing.assert_allclose(gg, oracle_ops.gather_points_grad(go, idx, n), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("b,c,n,np_,ns", [(2, 3, 5000, 300, 64), (2, 10, 4096, 256, 32), (1, 131, 600, 77, 16), (1, 1, 1, 1, 1)])
def test_group_and_grad(oracle_ops, b, c, n, np_, ns):
    rng = np.random.default_rng(1)
    pts = rng.standard_normal((b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, np_, ns)).astype(np.int32)
    got = ext.group_points(dev(pts), dev(idx)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.group_points(pts, idx))
    go = rng.standard_normal((b, c, np_, ns)).astype(np.float32)
    gg = ext.group_points_grad(dev(go), dev(idx), n).cpu().numpy()
    np.testing.assert_allclose(gg, oracle_ops.group_points_grad(go, idx, n), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("b,n,c,m,ns,r,norm", [(2, 3000, 7, 128, 16, 0.3, True), (2, 9000, 132, 200, 64, 0.2, True),
                                                  (1, 700, 1, 33, 8, 0.5, False), (3, 2048, 128, 77, 32, 0.4, True)])
def test_query_and_group_from_point_major_twin(b, n, c, m, ns, r, norm):
    """One-pass grouping + centring + cat from the point-major cloud == the reference sequence
    (two group_points, subtract, divide, cat; pointnet2_utils.py:347-359), bit for bit."""
    pc = synthetic.make_batch(b, n, c, first_scene=61).cuda()
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()
    _, new_xyz = ext.furthest_point_sampling(xyz, m, return_xyz=True)
    grouper = pu.QueryAndGroup(r, ns, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=norm)
    want, want_xyz = grouper(xyz, new_xyz, feats)
    view = pc[..., 3:].transpose(1, 2)               # what Pointnet2Backbone hands to SA1
    view._bqa_pm = pc[..., 3:]
    got, got_xyz = grouper(xyz, new_xyz, view)
    assert torch.equal(got, want) and torch.equal(got_xyz, want_xyz)
    # a gradient request must take the differentiable path
    f2 = feats.clone().requires_grad_(True)
    f2._bqa_pm = pc[..., 3:]
    out, _ = grouper(xyz, new_xyz, f2)
    out.sum().backward()
    assert f2.grad is not None and torch.equal(out, want)


def test_group_autograd_roundtrip():
    """grouping_operation backward == torch's own gather backward."""
    torch.manual_seed(0)
    f = torch.randn(2, 6, 500, device="cuda", requires_grad=True)
    idx = torch.randint(0, 500, (2, 40, 8), device="cuda", dtype=torch.int32)
    out = pu.grouping_operation(f, idx)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    f2 = f.detach().clone().requires_grad_(True)
    ref = torch.gather(f2.unsqueeze(2).expand(-1, -1, 40, -1), 3,
                       idx.long().unsqueeze(1).expand(-1, 6, -1, -1))
    assert torch.equal(out, ref)
    (ref * w).sum().backward()
    torch.testing.assert_close(f.grad, f2.grad, rtol=1e-5, atol=1e-5)


def test_transpose_to_point_major():
    x = torch.randn(3, 37, 1001, device="cuda")
    assert torch.equal(ext.transpose_to_point_major(x), x.transpose(1, 2).contiguous())


# ------------------------------------------------- three_nn / three_interpolate ---

@pytest.mark.parametrize("b,n,m", [(2, 512, 256), (2, 1024, 512), (3, 100, 3), (2, 50, 2), (2, 9, 1), (1, 3000, 1500)])
def test_three_nn_matches_oracle(oracle_ops, b, n, m):
    xyz = scenes(b, max(n, m) + 5, first=21)
    unknown, known = xyz[:, :n].copy(), xyz[:, 3:3 + m].copy()
    d2w, iw = oracle_ops.three_nn(unknown, known)
    d2, idx = ext.three_nn(dev(unknown), dev(known))
    np.testing.assert_array_equal(idx.cpu().numpy(), iw)
    np.testing.assert_array_equal(d2.cpu().numpy(), d2w)      # inf where m < 3
    dist, idx2 = pu.three_nn(dev(unknown), dev(known))
    np.testing.assert_array_equal(dist.cpu().numpy(), np.sqrt(d2w))


def test_three_nn_duplicate_known_points(oracle_ops):
    rng = np.random.default_rng(3)
    known = rng.uniform(-1, 1, (2, 40, 3)).astype(np.float32)[:, rng.integers(0, 40, 400)]
    unknown = rng.uniform(-1, 1, (2, 333, 3)).astype(np.float32)
    d2w, iw = oracle_ops.three_nn(unknown, known)
    d2, idx = ext.three_nn(dev(unknown), dev(known))
    np.testing.assert_array_equal(idx.cpu().numpy(), iw)


@pytest.mark.parametrize("b,c,m,n", [(2, 256, 256, 512), (2, 256, 512, 1024), (1, 7, 5, 33), (1, 1, 3, 1)])
def test_three_interpolate_and_grad(oracle_ops, b, c, m, n):
    rng = np.random.default_rng(4)
    feats = rng.standard_normal((b, c, m)).astype(np.float32)
    idx = rng.integers(0, m, (b, n, 3)).astype(np.int32)
    w = rng.uniform(0, 1, (b, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    got = ext.three_interpolate(dev(feats), dev(idx), dev(w)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_ops.three_interpolate(feats, idx, w))   # same fma chain
    go = rng.standard_normal((b, c, n)).astype(np.float32)
    gg = ext.three_interpolate_grad(dev(go), dev(idx), dev(w), m).cpu().numpy()
    np.testing.assert_allclose(gg, oracle_ops.three_interpolate_grad(go, idx, w, m), rtol=1e-4, atol=1e-4)


def test_three_interpolate_gradcheck_like_reference():
    """The reference's one unit test (pointnet2_test.py:18-30), same idx / weights / tolerances."""
    from torch.autograd import gradcheck
    feats = torch.randn(1, 2, 4, device="cuda", requires_grad=True)
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32, device="cuda")
    weight = torch.tensor([[[1, 1, 1], [2, 2, 2]]], dtype=torch.float32, device="cuda")
    assert gradcheck(lambda x: pu.three_interpolate(x, idx, weight), feats, atol=1e-1, rtol=1e-1)


def test_interp_vs_reference_extension(ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not built")
    xyz = dev(scenes(4, 1024, first=30))
    unknown, known = xyz[:, :1024].contiguous(), xyz[:, :512].contiguous()
    d2r, ir = ref_ext.three_nn(unknown, known)
    d2, idx = ext.three_nn(unknown, known)
    assert torch.equal(idx, ir) and torch.equal(d2, d2r)
    feats = torch.randn(4, 256, 512, device="cuda")
    w = torch.rand(4, 1024, 3, device="cuda")
    assert torch.equal(ext.three_interpolate(feats, idx, w), ref_ext.three_interpolate(feats, ir, w))
    pts = torch.randn(4, 10, 40000, device="cuda")
    gidx = torch.randint(0, 40000, (4, 2048, 64), device="cuda", dtype=torch.int32)
    assert torch.equal(ext.group_points(pts, gidx), ref_ext.group_points(pts, gidx))
    fidx = torch.randint(0, 40000, (4, 2048), device="cuda", dtype=torch.int32)
    assert torch.equal(ext.gather_points(pts, fidx), ref_ext.gather_points(pts, fidx))
    go = torch.randn(4, 10, 2048, 64, device="cuda")
    torch.testing.assert_close(ext.group_points_grad(go, gidx, 40000),
                               ref_ext.group_points_grad(go, gidx, 40000), rtol=1e-4, atol=1e-4)


# -------------------------------------------------------------- error behaviour ---

def test_errors_like_reference_checks():
    x = torch.randn(2, 100, 3)
    with pytest.raises(RuntimeError):      # CPU tensor: "CPU not supported" in the reference
        ext.furthest_point_sampling(x, 10)
    xc = torch.randn(2, 3, 100, device="cuda").transpose(1, 2)
    with pytest.raises(RuntimeError):      # CHECK_CONTIGUOUS
        ext.furthest_point_sampling(xc, 10)
    with pytest.raises(RuntimeError):      # CHECK_IS_INT
        ext.gather_points(torch.randn(1, 3, 10, device="cuda"), torch.zeros(1, 4, device="cuda", dtype=torch.int64))
    with pytest.raises(RuntimeError):      # CHECK_IS_FLOAT
        ext.ball_query(torch.zeros(1, 4, 3, device="cuda", dtype=torch.float64),
                       torch.zeros(1, 9, 3, device="cuda"), 0.1, 4)


def test_ops_run_on_current_stream_and_count_launches():
    from bridgeqa_b200 import _native
    x = dev(scenes(2, 4096))
    s = torch.cuda.Stream()
    before = _native.launch_count()
    with torch.cuda.stream(s):
        a = ext.furthest_point_sampling(x, 128)
    s.synchronize()
    assert _native.launch_count() == before + 1
    assert torch.equal(a, ext.furthest_point_sampling(x, 128))


# ------------------------------------------------------------------ slice forms ---

@pytest.mark.parametrize("b,n,m,cuts", [(3, 40000, 512, [1, 100, 101, 400, 512]), (2, 3000, 300, [1, 150, 300]),
                                        (2, 20000, 256, [1, 64, 128, 192, 256])])
def test_fps_slices_equal_one_call(b, n, m, cuts):
    """Sliced (resumable) sampling produces exactly the indices / centres of one full call."""
    import ctypes
    from bridgeqa_b200 import _native as N
    xyz = dev(scenes(b, n, first=13))
    want, want_xyz = ext.furthest_point_sampling(xyz, m, return_xyz=True)
    inds = torch.full((b, m), -1, dtype=torch.int32, device="cuda")
    new_xyz = torch.zeros((b, m, 3), device="cuda")
    state = torch.empty((b, n), device="cuda")
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        N.call("bqa_furthest_point_sampling_slice", b, n, m, lo, hi, N.ptr(xyz), N.ptr(inds), N.ptr(new_xyz),
               N.ptr(state), 1, N.stream_ptr(xyz.device))
    assert torch.equal(inds, want) and torch.equal(new_xyz, want_xyz)


def test_ball_query_slices_equal_one_call():
    import ctypes
    from bridgeqa_b200 import _native as N
    b, n, m, ns, r = 2, 20000, 512, 32, 0.25
    xyz = dev(scenes(b, n, first=17))
    _, centres = ext.furthest_point_sampling(xyz, m, return_xyz=True)
    want = ext.ball_query(centres, xyz, r, ns)
    idx = torch.full((b, m, ns), -1, dtype=torch.int32, device="cuda")
    for lo, cnt in [(0, 128), (128, 256), (384, 128)]:
        nbytes = N.lib().bqa_ball_query_workspace_bytes(b, n, cnt, ns)
        work = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device="cuda")
        N.call("bqa_ball_query_slice", b, n, m, lo, cnt, ctypes.c_float(r), ns, N.ptr(centres), N.ptr(xyz),
               N.ptr(idx), N.ptr(work) if nbytes else ctypes.c_void_p(0), N.stream_ptr(xyz.device))
    assert torch.equal(idx, want)


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,c", [(2, 1000, 7), (1, 257, 7), (3, 256, 1), (2, 513, 13), (2, 300, 29), (2, 300, 132)])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_point_major_16_of_the_input_cloud(b, n, c, prec):
    """The 16-bit point-major twin of the backbone's input features (the SA1 gather source), converted row-wise
    from the (B, N, 3 + C) cloud, one 16-byte chunk per thread (narrow and wide rows, ragged row counts, zero
    padding of the last chunk).  Bit-equal to torch's own round-to-nearest conversion."""
    from bridgeqa_b200 import fused
    old = fused.precision()
    fused.set_precision(prec)
    try:
        g = torch.Generator().manual_seed(c * 1000 + n)
        cloud = (torch.randn(b, n, 3 + c, generator=g) * 3).cuda()
        feats = cloud[..., 3:].transpose(1, 2)                 # (B, C, N) view, as the backbone makes it
        feats._bqa_cloud = cloud
        twin = fused.point_major_16(feats)
        stride = (c + 7) // 8 * 8
        assert twin.shape == (b, n, stride) and twin.dtype == torch.int16
        dt = torch.float16 if prec == "fp16" else torch.bfloat16
        want = torch.zeros(b, n, stride, dtype=dt, device="cuda")
        want[..., :c] = cloud[..., 3:].to(dt)
        assert torch.equal(twin, want.view(torch.int16))
        # and the transpose-convert kernel used when the features are a tensor of their own
        twin2 = fused.point_major_16(feats.contiguous())
        assert torch.equal(twin2, want.view(torch.int16))
    finally:
        fused.set_precision(old)
