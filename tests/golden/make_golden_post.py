"""Generate tests/golden/ref_postprocess.npz by running the UNMODIFIED reference utilities

    utils/box_util.py:302-325   get_3d_box_batch          (box decode)
    utils/nms.py:74-152         nms_3d_faster, nms_3d_faster_samecls   

from /root/reference on the CPU (pure NumPy; only runnable in the build container).

    python tests/golden/make_golden_post.py

Third-party modules the reference imports at module level but that are absent here (plyfile,
trimesh, matplotlib -- I/O / plotting helpers, not on this path) are stubbed in sys.modules so `utils.nms` imports.
`param2obb_batch` lives in the un-vendored data/scannet/model_util_scannet.py (dangling symlink,
SURVEY 8c): its published VoteNet/ScanRefer form is restated below (size = mean_size[class] +
residual; heading = -class2angle_batch; ScanNet: class2angle_batch == 0; "bins": class * 2pi/nh +
residual wrapped to (-pi, pi]).

NMS ties: the reference walks np.argsort(score) from its end.  NumPy's default argsort does not
define the order of equal scores (here: numpy %s with the AVX-512 sort, which is not stable even for
8 elements), so for the tie cases two reference runs are stored: `*_default` with the argsort of
THIS machine, and `*_stable` with np.argsort forced to kind="stable" inside the same unmodified
function.  The device kernel implements the stable order; tie-free cases have one answer.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _import_reference():
    for name in ("plyfile", "trimesh", "matplotlib", "matplotlib.pyplot"):
        try:
            __import__(name)
        except ImportError:
            m = types.ModuleType(name)
            m.PlyData = m.PlyElement = object
            m.cm = types.SimpleNamespace(jet=None)     # default argument at pc_utils.py:198
            sys.modules[name] = m
    sys.path.insert(0, REF)
    import utils.box_util as box_util
    import utils.nms as nms
    return box_util, nms


def param2obb_batch(center, heading_class, heading_residual, size_class, size_residual, mean_size_arr,
                    heading_mode, num_heading_bin):
    """Restatement of ScannetDatasetConfig.param2obb_batch (un-vendored; VoteNet model_util_scannet.py)."""
    if heading_mode == "zero":
        heading_angle = np.zeros(heading_class.shape[0])
    else:
        heading_angle = heading_class * (2 * np.pi / float(num_heading_bin)) + heading_residual
        heading_angle[heading_angle > np.pi] -= 2 * np.pi
    box_size = mean_size_arr[size_class, :] + size_residual
    obb = np.zeros((heading_class.shape[0], 7))
    obb[:, 0:3] = center
    obb[:, 3:6] = box_size
    obb[:, 6] = heading_angle * -1
    return obb


def random_boxes(rng, k, with_cls, ties):
    c = rng.uniform(-3, 3, size=(k, 3))
    s = rng.uniform(0.2, 1.5, size=(k, 3))
    if ties == "none":
        score = (rng.permutation(k).astype(np.float64) + 0.5) / k
    elif ties == "saturated":       # softmax probabilities that round to exactly 1.0f / a few levels
        score = rng.choice(np.array([1.0, 1.0, 0.99999994, 0.5, 0.25], dtype=np.float32), size=k).astype(np.float64)
    else:                           # every score shared by several boxes
        score = rng.randint(0, max(2, k // 6), size=k).astype(np.float64) / max(2, k // 6)
    cols = [c - s / 2, c + s / 2, score[:, None]]
    if with_cls:
        cols.append(rng.randint(0, 4, size=(k, 1)).astype(np.float64))
    # float32 values (what the device receives), stored as the float64 array the reference builds
    return np.concatenate(cols, 1).astype(np.float32).astype(np.float64)


def main():
    box_util, nms = _import_reference()
    rng = np.random.RandomState(20261017)
    out = {}

    # ---- box decode -----------------------------------------------------------------------
    cases = []
    for ci, (mode, nh, k) in enumerate([("zero", 1, 256), ("bins", 12, 256), ("bins", 1, 64)]):
        ns = 18
        mean_size = rng.uniform(0.3, 2.0, size=(ns, 3)).astype(np.float32)
        center = rng.uniform(-4, 4, size=(k, 3)).astype(np.float32)
        heading_scores = rng.normal(size=(k, nh)).astype(np.float32)
        heading_res_norm = rng.uniform(-1, 1, size=(k, nh)).astype(np.float32)
        size_scores = rng.normal(size=(k, ns)).astype(np.float32)
        size_res_norm = rng.uniform(-0.3, 0.3, size=(k, ns, 3)).astype(np.float32)
        # proposal_module.py:131-136 (float32 tensors)
        heading_residuals = heading_res_norm * np.float32(np.pi / nh)
        size_residuals = size_res_norm * mean_size[None]
        # proposal_module.py:87-104
        hcls = heading_scores.argmax(-1)
        hres = np.take_along_axis(heading_residuals, hcls[:, None], 1)[:, 0]
        scls = size_scores.argmax(-1)
        sres = size_residuals[np.arange(k), scls]
        obb = param2obb_batch(center, hcls, hres, scls, sres, mean_size, mode, nh)
        corners = box_util.get_3d_box_batch(obb[:, 3:6], obb[:, 6], obb[:, 0:3])      # the reference itself
        p = "decode%d_" % ci
        out[p + "mode"] = np.array(mode)
        out[p + "mean_size"], out[p + "center"] = mean_size, center
        out[p + "heading_scores"], out[p + "heading_residuals_normalized"] = heading_scores, heading_res_norm
        out[p + "size_scores"], out[p + "size_residuals_normalized"] = size_scores, size_res_norm
        out[p + "bbox_corner"] = corners
        cases.append(ci)
    out["decode_cases"] = np.array(cases)

    # ---- NMS ------------------------------------------------------------------------------
    real_argsort = np.argsort
    ncase = 0
    for k, thr, old, cls, ties in [(256, 0.25, False, False, "none"), (256, 0.25, True, False, "none"),
                                   (256, 0.1, False, True, "none"), (37, 0.5, False, False, "none"),
                                   (1000, 0.25, False, True, "none"), (1, 0.25, False, False, "none"),
                                   (256, 0.25, False, False, "saturated"), (256, 0.25, False, True, "saturated"),
                                   (200, 0.25, True, False, "many"), (16, 0.25, False, False, "many"),
                                   (600, 0.1, False, True, "many")]:
        boxes = random_boxes(rng, k, cls, ties)
        fn = nms.nms_3d_faster_samecls if cls else nms.nms_3d_faster
        default = np.asarray(fn(boxes, thr, old), dtype=np.int64)
        np.argsort = lambda a, *args, **kw: real_argsort(a, kind="stable")
        nms.np.argsort = np.argsort
        try:
            stable = np.asarray(fn(boxes, thr, old), dtype=np.int64)
        finally:
            np.argsort = real_argsort
            nms.np.argsort = real_argsort
        p = "nms%d_" % ncase
        out[p + "boxes"] = boxes.astype(np.float32)
        out[p + "thr"], out[p + "old"], out[p + "cls"] = np.float64(thr), np.int64(old), np.int64(cls)
        out[p + "ties"] = np.array(ties)
        out[p + "pick_default"], out[p + "pick_stable"] = default, stable
        if ties == "none":
            assert np.array_equal(default, stable)
        ncase += 1
    out["nms_cases"] = np.int64(ncase)
    out["meta_numpy"] = np.array(np.__version__)
    path = os.path.join(ROOT, "tests", "golden", "ref_postprocess.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for i in range(ncase):
        print(i, str(out["nms%d_ties" % i]), len(out["nms%d_pick_default" % i]),
              "default==stable:", np.array_equal(out["nms%d_pick_default" % i], out["nms%d_pick_stable" % i]))


if __name__ == "__main__":
    __doc__ = __doc__ % np.__version__
    main()
