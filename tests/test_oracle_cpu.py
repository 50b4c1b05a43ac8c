"""CPU-only checks of the oracle (oracle/pointnet2_oracle.c + oracle/modules_cpu.py):
against the committed golden vectors (outputs of the reference's CUDA extension and of the
reference's Python layer) and against independent brute-force / literal re-simulations."""
import glob
import os

import numpy as np
import pytest
import torch

from bridgeqa_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def scenes(b, n, first=0):
    return synthetic.make_batch(b, n, 0, first_scene=first)[..., :3].contiguous().numpy()


# ---- literal simulation of the reference FPS kernel (threads + tree), tiny sizes only ----

def fps_literal(xyz, m, block_size):
    """Pure-Python transliteration of sampling_gpu.cu:69-173 with `block_size` threads,
    float32 arithmetic with an exact fma emulated in float64 (products of float32 are exact
    in float64; one rounding per fma step)."""
    f32 = np.float32

    def fma(a, b, c):
        return f32(np.float64(a) * np.float64(b) + np.float64(c))

    n = xyz.shape[0]
    temp = np.full(n, 1e10, dtype=np.float32)
    idxs = np.zeros(m, dtype=np.int32)
    old = 0
    for j in range(1, m):
        dists = np.full(block_size, -1.0, dtype=np.float32)
        dists_i = np.zeros(block_size, dtype=np.int64)
        x1, y1, z1 = xyz[old]
        for tid in range(block_size):
            best, besti = f32(-1), 0
            for k in range(tid, n, block_size):
                x2, y2, z2 = xyz[k]
                mag = fma(z2, z2, fma(x2, x2, f32(y2 * y2)))
                if np.float64(mag) <= 1e-3:
                    continue
                dx, dy, dz = f32(x2 - x1), f32(y2 - y1), f32(z2 - z1)
                d = fma(dz, dz, fma(dx, dx, f32(dy * dy)))
                d2 = min(d, temp[k])
                temp[k] = d2
                if d2 > best:
                    besti, best = k, d2
            dists[tid], dists_i[tid] = best, besti
        s = block_size // 2
        while s >= 1:
            for tid in range(s):
                v1, v2 = dists[tid], dists[tid + s]
                i1, i2 = dists_i[tid], dists_i[tid + s]
                dists[tid] = max(v1, v2)
                dists_i[tid] = i2 if v2 > v1 else i1
            s //= 2
        old = int(dists_i[0])
        idxs[j] = old
    return idxs


@pytest.mark.parametrize("n,m", [(40, 12), (130, 20), (600, 24)])
def test_fps_oracle_equals_literal_kernel_simulation(oracle_ops, n, m):
    rng = np.random.default_rng(n)
    xyz = rng.integers(-3, 4, (n, 3)).astype(np.float32) * np.float32(0.5)   # lattice: many ties
    xyz[::9] *= np.float32(0.01)                                             # skipped points
    bs = oracle_ops.opt_n_threads(n)
    want = fps_literal(xyz, m, bs)
    got = oracle_ops.furthest_point_sampling(xyz[None], m)[0]
    np.testing.assert_array_equal(got, want)


def _cell_order(xyz, h):
    """A cell-sorted order like ball_query_grid.cu's counting sort (any spatially coherent order
    exercises the rule; the exact cell function does not matter for soundness)."""
    c = np.floor((xyz - xyz.min(0)) / np.float32(h)).astype(np.int64)
    key = (c[:, 2] * 4096 + c[:, 1]) * 4096 + c[:, 0]
    return np.argsort(key, kind="stable")


@pytest.mark.parametrize("n,m,run,scale", [(8192, 300, 448, 1.0), (8192, 300, 576, 1.0), (6000, 200, 64, 1.0),
                                           (8192, 200, 576, 1000.0), (4096, 150, 32, 1e-2)])
def test_fps_pruning_rule_is_sound(oracle_ops, n, m, run, scale):
    """The skip rule of fps_sorted.cu, replayed on the CPU: a warp holding a run of cell-sorted points
    with bounding box [lo, hi] skips an iteration when  lb * 0.99999f >= wmax  (lb = fp32 squared
    distance from the new sample to the box, wmax = max running min-distance of the run).  Sound
    means: in every skipped (iteration, run) NO point's min-distance would have changed, i.e. the
    fp32 distance the reference computes is >= the stored one for every point of the run.  Checked
    against the full update with the reference's arithmetic (fma emulated exactly), on duplicates,
    skip-set points, large and tiny coordinate scales; the samples come from the C oracle."""
    f32 = np.float32
    xyz = (scenes(1, n, first=17)[0] * f32(scale)).astype(np.float32)
    xyz[5:40] = xyz[5]                                        # duplicates (zero-size boxes inside runs)
    xyz[100:108] = f32(0.004) * f32(min(scale, 1.0))          # |p|^2 <= 1e-3: never updated, never picked
    inds = oracle_ops.furthest_point_sampling(xyz[None], m)[0]
    order = _cell_order(xyz, 0.2 * scale)
    pts = xyz[order]
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)                                 # original index -> position in the sorted order
    x, y, z = (pts[:, i].astype(np.float64) for i in range(3))
    mag = (z * z + (x * x + (y * y).astype(f32).astype(np.float64)).astype(f32).astype(np.float64)).astype(f32)
    selectable = ~(mag.astype(np.float64) <= 1e-3)
    td = np.where(selectable, f32(1e10), f32(-np.inf)).astype(np.float32)
    nrun = (n + run - 1) // run
    pad = nrun * run - n
    def runs(a, fill):
        return np.concatenate([a, np.full(pad, fill, a.dtype)]).reshape(nrun, run)
    lo = np.stack([runs(pts[:, i], f32(np.inf)).min(1) for i in range(3)], 1)
    hi = np.stack([runs(pts[:, i], f32(-np.inf)).max(1) for i in range(3)], 1)
    wmax = np.full(nrun, np.inf, dtype=np.float32)            # inf: not evaluated yet
    skipped = total = 0
    for j in range(1, m):
        s = xyz[inds[j - 1]]
        dx = (pts[:, 0] - s[0]).astype(np.float64)
        dy = (pts[:, 1] - s[1]).astype(np.float64)
        dz = (pts[:, 2] - s[2]).astype(np.float64)
        d = (dz * dz + (dx * dx + (dy * dy).astype(f32).astype(np.float64)).astype(f32).astype(np.float64)).astype(f32)
        would_change = selectable & (d < td)                  # fminf(d, td) != td
        e = np.maximum(np.maximum(lo - s, s - hi), f32(0)).astype(np.float32)
        lb = (e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2]).astype(np.float32)
        skip = lb * f32(0.99999) >= wmax
        changed_runs = runs(would_change, False).any(1)
        assert not (skip & changed_runs).any(), "iteration %d: a skipped run had a point to update" % j
        td = np.where(would_change, d, td)
        # runs that did the update refresh their max (the kernel's wmax); -inf when nothing is selectable
        new_max = runs(td, f32(-np.inf)).max(1)
        wmax = np.where(skip, wmax, new_max).astype(np.float32)
        skipped += int(skip.sum())
        total += nrun
        # the replay's arithmetic is the oracle's: its next sample holds the maximal min-distance
        assert td[pos[inds[j]]] == td.max(), j
    assert skipped > 0.3 * total, "the rule was hardly exercised (%d of %d)" % (skipped, total)


def _prefix_proof_flag(xyz, m):
    """CPU statement of fps_prefix_v_kernel + fps_prefix_check_kernel (csrc/fps.cu): 0 = the sampling of
    this cloud is PROVABLY the identity prefix, 1 = not provable."""
    f32 = np.float32
    n = xyz.shape[0]
    p = xyz.astype(np.float64)

    def sq(a, b):                                             # fma(dz,dz,fma(dx,dx,dy*dy)), exact emulation
        d = a - b
        return (d[..., 2] * d[..., 2]
                + (d[..., 0] * d[..., 0] + (d[..., 1] * d[..., 1]).astype(f32).astype(np.float64)).astype(f32)
                .astype(np.float64)).astype(f32)

    mag = (p[:, 2] * p[:, 2] + (p[:, 0] * p[:, 0] + (p[:, 1] * p[:, 1]).astype(f32).astype(np.float64))
           .astype(f32).astype(np.float64)).astype(f32)
    if (mag[1:].astype(np.float64) <= 1e-3).any():
        return 1
    r = np.full(n, 1e10, dtype=np.float32)                    # R(k, j) for the current j
    for j in range(1, m):
        r = np.minimum(r, sq(p, p[j - 1]))                    # now min over i < j
        v = r[j]                                              # V[j]
        if not v > 0:
            return 1
        if j + 1 < n and not (r[j + 1:] < v).all():
            return 1
    return 0


def test_fps_prefix_proof_is_sound(oracle_ops):
    """The parallel proof that replaces the SA2-4 sampling chains: whenever the proof
    accepts a cloud, the oracle's furthest point sampling of every prefix of it to every m' <= m IS
    the identity prefix; generic sampled clouds are accepted (the test is not vacuous), skip-set points
    are always rejected, tie-ridden lattices are rejected or still identity."""
    base = scenes(1, 6000, first=23)[0]
    order = oracle_ops.furthest_point_sampling(base[None], 1024)[0]
    cloud = base[order]                                        # a cloud in sampling order, like SA2's input
    m = 512
    assert _prefix_proof_flag(cloud, m) == 0
    for n2, m2 in [(1024, 512), (1024, 100), (700, 512), (512, 512), (513, 256)]:
        got = oracle_ops.furthest_point_sampling(cloud[None, :n2], m2)[0]
        np.testing.assert_array_equal(got, np.arange(m2))
    bad = cloud.copy()
    bad[300] = [0.01, 0.01, 0.02]                              # |p|^2 <= 1e-3
    assert _prefix_proof_flag(bad, m) == 1
    # a far-away point late in the cloud breaks the order: must be rejected
    far = cloud.copy()
    far[900] = far[0] + np.float32(50.0)
    assert _prefix_proof_flag(far, m) == 1
    assert not np.array_equal(oracle_ops.furthest_point_sampling(far[None], m)[0], np.arange(m))
    # lattices (exact ties everywhere): whatever the proof says must be true
    rng = np.random.default_rng(3)
    for trial in range(6):
        g = rng.integers(0, 6, (400, 3)).astype(np.float32) * np.float32(0.5) + np.float32(1.0)
        g = np.unique(g, axis=0)
        rng.shuffle(g)
        o = oracle_ops.furthest_point_sampling(g[None], len(g))[0]
        lat = g[o]
        mm = min(64, len(lat))
        if _prefix_proof_flag(lat, mm) == 0:
            np.testing.assert_array_equal(oracle_ops.furthest_point_sampling(lat[None], mm)[0], np.arange(mm))


def test_opt_n_threads_rule(oracle_ops):
    for n, want in [(1, 1), (2, 2), (3, 2), (511, 256), (512, 512), (513, 512), (1024, 512),
                    (2048, 512), (20000, 512), (40000, 512), (100000, 512)]:
        assert oracle_ops.opt_n_threads(n) == want


def test_fps_tie_break_is_bitrev_order(oracle_ops):
    """All points equidistant from point 0 except point 0 itself: the second pick is the tied
    point minimising (bitrev9(k mod 512), k) -- SURVEY.md 8a rule 5."""
    n = 1500
    ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
    xyz = np.zeros((1, n, 3), dtype=np.float32)
    # points on a lattice sphere would not be exactly equidistant; use two values only
    xyz[0, :, 0] = 1.0
    xyz[0, 0] = (3.0, 0.0, 0.0)            # seed; everyone else at squared distance 4
    got = oracle_ops.furthest_point_sampling(xyz, 2)[0]

    def bitrev9(v):
        return int("{:09b}".format(v)[::-1], 2)
    cand = min(range(1, n), key=lambda k: (bitrev9(k % 512), k))
    assert got[1] == cand


def test_ball_query_bruteforce(oracle_ops):
    xyz = scenes(2, 3000, first=4)
    centres = xyz[:, ::30].copy()
    centres[:, -3:] += 50.0                       # empty balls
    r, ns = 0.35, 16
    got = oracle_ops.ball_query(centres, xyz, r, ns)
    r2 = np.float32(r) * np.float32(r)
    for b in range(2):
        for j in range(0, centres.shape[1], 7):
            d = (centres[b, j].astype(np.float64) - xyz[b].astype(np.float64))
            d2 = (d ** 2).sum(-1)
            # exclude borderline cases from the float64 brute force
            hits = np.nonzero(d2 < r2)[0]
            if np.any(np.abs(d2 - r2) < 1e-6):
                continue
            row = np.zeros(ns, np.int32)
            if len(hits):
                row[:] = hits[0]
                row[:min(ns, len(hits))] = hits[:ns]
            np.testing.assert_array_equal(got[b, j], row)
    assert (got[:, -3:] == 0).all()


def test_three_nn_bruteforce(oracle_ops):
    xyz = scenes(2, 900, first=6)
    unknown, known = xyz[:, :500], xyz[:, 500:]
    d2, idx = oracle_ops.three_nn(unknown, known)
    full = ((unknown[:, :, None].astype(np.float64) - known[:, None].astype(np.float64)) ** 2).sum(-1)
    order = np.argsort(full, axis=-1, kind="stable")[..., :3]
    np.testing.assert_allclose(d2, np.take_along_axis(full, order, -1), rtol=1e-5, atol=1e-7)
    assert (idx == order).mean() > 0.999          # float32 vs float64 near-ties only
    d2s, idxs = oracle_ops.three_nn(unknown, known[:, :2])
    assert np.isinf(d2s[..., 2]).all() and (idxs[..., 2] == 0).all()


def test_gather_group_interpolate_numpy(oracle_ops):
    rng = np.random.default_rng(2)
    pts = rng.standard_normal((2, 5, 300)).astype(np.float32)
    idx = rng.integers(0, 300, (2, 40, 8)).astype(np.int32)
    np.testing.assert_array_equal(oracle_ops.group_points(pts, idx),
                                  np.stack([pts[b][:, idx[b]] for b in range(2)]))
    g = oracle_ops.gather_points(pts, idx[:, :, 0].copy())
    np.testing.assert_array_equal(g, np.stack([pts[b][:, idx[b, :, 0]] for b in range(2)]))
    go = rng.standard_normal((2, 5, 40, 8)).astype(np.float32)
    want = np.zeros((2, 5, 300), np.float64)
    for b in range(2):
        for c in range(5):
            np.add.at(want[b, c], idx[b].reshape(-1), go[b, c].reshape(-1))
    np.testing.assert_allclose(oracle_ops.group_points_grad(go, idx, 300), want, rtol=1e-5, atol=1e-5)
    # the reference's own unit test setting (pointnet2_test.py:18-30)
    feats = rng.standard_normal((1, 2, 4)).astype(np.float32)
    i3 = np.array([[[0, 1, 2], [1, 2, 3]]], np.int32)
    w3 = np.array([[[1, 1, 1], [2, 2, 2]]], np.float32)
    out = oracle_ops.three_interpolate(feats, i3, w3)
    np.testing.assert_allclose(out[0, :, 0], feats[0, :, :3].sum(-1), rtol=1e-6)
    np.testing.assert_allclose(out[0, :, 1], 2 * feats[0, :, 1:].sum(-1), rtol=1e-6)
    gi = oracle_ops.three_interpolate_grad(np.ones((1, 2, 2), np.float32), i3, w3, 4)
    np.testing.assert_allclose(gi[0, 0], [1, 3, 3, 2])


# ---- golden vectors ---------------------------------------------------------------------

def test_oracle_modules_match_reference_python_layer():
    """oracle/modules_cpu.py vs outputs of the reference's own pointnet2_modules /
    backbone_module / voting_module (tests/golden/make_golden_cpu.py)."""
    from oracle import modules_cpu
    from bridgeqa_b200 import detector
    g = np.load(os.path.join(GOLDEN, "ref_python_layer.npz"))
    B, N, C = int(g["meta_B"]), int(g["meta_N"]), int(g["meta_C"])
    pc = synthetic.make_batch(B, N, C, first_scene=int(g["meta_first_scene"]))
    net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=C),
                                    seed=int(g["meta_backbone_seed"]))
    assert sorted(net.state_dict().keys()) == list(g["backbone_keys"])
    out = modules_cpu.backbone(pc.numpy(), net.state_dict())
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        np.testing.assert_array_equal(out[k], g[k], err_msg=k)
    for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz"):
        np.testing.assert_array_equal(out[k], g[k], err_msg=k)

    def sample(a, k=4096):
        flat = np.ascontiguousarray(a).reshape(-1)
        return flat[::max(1, flat.size // k)][:k]

    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        np.testing.assert_allclose(sample(out[k]), g[k + "_sample"], rtol=1e-4, atol=1e-5, err_msg=k)
        np.testing.assert_allclose(np.abs(out[k].astype(np.float64)).sum(), g[k + "_abssum"], rtol=1e-5)
    vote = synthetic.fill_state_dict(detector.VotingModule(1, 256), seed=int(g["meta_voting_seed"]))
    assert sorted(vote.state_dict().keys()) == list(g["voting_keys"])
    vxyz, vfeat = modules_cpu.voting(out["fp2_xyz"], out["fp2_features"], vote.state_dict(), "")
    np.testing.assert_allclose(vxyz, g["vote_xyz"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(sample(vfeat), g["vote_features_sample"], rtol=1e-4, atol=1e-5)


def test_oracle_matches_reference_extension_golden(oracle_ops):
    """tests/golden/ref_ext_*.npz hold outputs of the reference's CUDA extension run on the
    B200 box (tests/golden/make_golden_gpu.py); the C oracle must reproduce them bit for bit."""
    files = sorted(glob.glob(os.path.join(GOLDEN, "ref_ext_*.npz")))
    if not files:
        pytest.skip("no reference-extension golden files committed yet")
    for path in files:
        g = np.load(path)
        xyz = scenes(int(g["B"]), int(g["N"]), first=int(g["first_scene"]))
        inds = oracle_ops.furthest_point_sampling(xyz, int(g["npoint"]))
        np.testing.assert_array_equal(inds, g["fps_inds"], err_msg=path)
        centres = np.take_along_axis(xyz, inds[..., None].astype(np.int64), 1)
        bq = oracle_ops.ball_query(centres, xyz, float(g["radius"]), int(g["nsample"]))
        np.testing.assert_array_equal(bq, g["ball_idx"], err_msg=path)
        m = int(g["nn_m"])
        d2, idx = oracle_ops.three_nn(centres, np.ascontiguousarray(centres[:, :m]))
        np.testing.assert_array_equal(idx, g["nn_idx"], err_msg=path)
        np.testing.assert_array_equal(d2, g["nn_dist2"], err_msg=path)
        w = (1.0 / (np.sqrt(d2) + np.float32(1e-8))).astype(np.float32)
        w = (w / w.sum(-1, keepdims=True)).astype(np.float32)
        feats = np.ascontiguousarray(xyz[:, :m].transpose(0, 2, 1))
        np.testing.assert_array_equal(oracle_ops.three_interpolate(feats, idx, g["nn_weight"]),
                                      g["interp"], err_msg=path)
        grouped = oracle_ops.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), bq)
        np.testing.assert_array_equal(grouped[:, :, ::16], g["grouped_xyz_strided"], err_msg=path)
