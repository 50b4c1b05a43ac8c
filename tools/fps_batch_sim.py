"""Developer tool (CPU): how many samples per round could the one-SM sampling kernel certify at once?

Round: take the W best run records (stale upper bounds are exact at round start); candidate c_1 is the next
sample; c_i (i >= 2) is certified iff for every accepted c_j (j < i)
  (1) c_j cannot change run(c_i) at all (the kernels' box rule: lb * 0.99999 >= the run's max), and
  (2) the best of run(c_j) AFTER applying c_j to it (computed speculatively, read-only) is below c_i.
Every other run's record is an upper bound that already ranks below c_i.  The accepted prefix is then applied
in one pass.  Prints samples per round; checks the sequence against the oracle.

    python tools/fps_batch_sim.py [n] [npoint] [W]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bridgeqa_b200 import synthetic
from oracle import cpu_ops
from fps_prune_sim import cell_order


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    run = 64
    pc = synthetic.make_batch(1, n, 0, first_scene=3)
    xyz = pc[0, :, :3].numpy().copy()
    seq = cpu_ops.furthest_point_sampling(xyz[None], m)[0]
    order, g, h = cell_order(xyz)
    p = xyz[order].astype(np.float32)
    orig = order
    nr = (n + run - 1) // run
    pad = nr * run - n
    pp = np.concatenate([p, np.repeat(p[-1:], pad, 0)]) if pad else p
    valid = np.arange(nr * run) < n
    pr = pp.reshape(nr, run, 3)
    vr = valid.reshape(nr, run)
    lo, hi = pr.min(1), pr.max(1)
    mag = (pr ** 2).sum(2)
    td = np.where(vr & (mag > 1e-3), np.float32(1e10), np.float32(-np.inf)).astype(np.float32)

    def dist(run_ids, s):
        return ((pr[run_ids] - s) ** 2).sum(-1).astype(np.float32)

    got = [0]
    s0 = xyz[0]
    td = np.minimum(td, dist(np.arange(nr), s0)) if True else td
    td[~(vr & (mag > 1e-3))] = -np.inf
    rounds = []
    stops = {"box": 0, "own": 0, "window": 0}
    while len(got) < m:
        rec = td.max(1)
        top = np.argsort(-rec, kind="stable")[:W]
        cand = []
        for r in top:
            k = int(td[r].argmax())
            cand.append((r, k, rec[r], pr[r, k]))
        acc = [cand[0]]
        U = []
        for i in range(1, len(cand)):
            rj, kj, vj, pj = acc[-1]
            U.append(np.minimum(td[rj], dist(rj, pj)).max())
            ri, ki, vi, pi = cand[i]
            ok = True
            for (rj, kj, vj, pj), uj in zip(acc, U):
                e = np.maximum(np.maximum(lo[ri] - pj, pj - hi[ri]), 0)
                lb = np.float32((e * e).sum())
                if not (lb * np.float32(0.99999) >= vi):
                    ok = False; stops["box"] += 1; break
                if not (uj < vi):
                    ok = False; stops["own"] += 1; break
            if not ok:
                break
            acc.append(cand[i])
        else:
            stops["window"] += 1
        acc = acc[:m - len(got)]
        for r, k, v, pt in acc:
            got.append(int(orig[r * run + k]))
            td = np.minimum(td, dist(np.arange(nr), pt))
        rounds.append(len(acc))
    rounds = np.array(rounds)
    # exact duplicates tie (the kernels break ties by the reference's key, this replay by position): compare coordinates
    same = np.array_equal(xyz[np.array(got[:m])], xyz[seq[:m]])
    print("n=%d m=%d W=%d run=%d: %d rounds, %.2f samples per round (first 64 samples: %d rounds); histogram %s; "
          "stops %s; sequence == oracle: %s"
          % (n, m, W, run, len(rounds), (m - 1) / len(rounds), int(np.searchsorted(np.cumsum(rounds), 64)) + 1,
             np.bincount(rounds, minlength=W + 1).tolist(), stops, same))


if __name__ == "__main__":
    main()
