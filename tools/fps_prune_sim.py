"""Developer tool (CPU): how many points does the pruned sampling touch per iteration, as a function of
the run length and of the order the scene is sorted in?  Replays the oracle's sample sequence with the
kernels' box rule (skip a run when lb * 0.99999 >= the run's current max min-distance).

    python tools/fps_prune_sim.py [n] [npoint]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bridgeqa_b200 import synthetic
from oracle import cpu_ops


def cell_order(xyz, r=0.2, morton=False, kcells=32768):
    mn = xyz.min(0)
    ext = xyz.max(0) - mn
    h = r
    while True:
        g = np.minimum(ext / h, 65535).astype(np.int64) + 1
        if g.prod() <= kcells:
            break
        h *= 1.25
    c = np.clip(np.floor((xyz - mn) / h).astype(np.int64), 0, g - 1)
    if morton:
        def spread(v):
            o = np.zeros_like(v)
            for b in range(10):
                o |= ((v >> b) & 1) << (3 * b)
            return o
        key = spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)
    else:
        key = (c[:, 2] * g[1] + c[:, 1]) * g[0] + c[:, 0]
    return np.argsort(key, kind="stable"), g, h


def simulate(xyz, seq, order, run):
    p = xyz[order].astype(np.float32)
    n = len(p)
    nr = (n + run - 1) // run
    pad = nr * run - n
    pp = np.concatenate([p, np.repeat(p[-1:], pad, 0)]) if pad else p
    pr = pp.reshape(nr, run, 3)
    lo, hi = pr.min(1), pr.max(1)
    td = np.full((nr, run), 1e10, np.float32)
    rmax = np.full(nr, np.inf, np.float32)
    active_pts = []
    for j in range(1, len(seq)):
        s = xyz[seq[j - 1]]
        e = np.maximum(np.maximum(lo - s, s - hi), 0)
        lb = (e * e).sum(1)
        act = ~(lb * np.float32(0.99999) >= rmax)
        idx = np.nonzero(act)[0]
        d = ((pr[idx] - s) ** 2).sum(2)
        td[idx] = np.minimum(td[idx], d)
        rmax[idx] = td[idx].max(1)
        active_pts.append(len(idx) * run)
    return np.array(active_pts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    pc = synthetic.make_batch(1, n, 0, first_scene=3)
    xyz = pc[0, :, :3].numpy().copy()
    seq = cpu_ops.furthest_point_sampling(xyz[None], m)[0]
    for morton in (False, True):
        order, g, h = cell_order(xyz, morton=morton)
        for run in (32, 64, 128, 256, 576):
            a = simulate(xyz, seq, order, run)
            print("morton=%d grid=%s h=%.3f run=%4d: mean active points/iter %7.0f (%.2f %%), median %6.0f, "
                  "iters 1-64 mean %7.0f, last 1024 mean %6.0f, total %.2fM"
                  % (morton, tuple(int(v) for v in g), h, run, a.mean(), 100 * a.mean() / n, np.median(a),
                     a[:64].mean(), a[-1024:].mean(), a.sum() / 1e6))


if __name__ == "__main__":
    main()
