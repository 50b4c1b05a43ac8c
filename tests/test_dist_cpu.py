"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: scene sharding, the flat-bucket
gradient all-reduce and the max-over-ranks timing reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bridgeqa_b200 import detector, distributed as D, synthetic
    try:
        # every rank builds the same model, a rank-dependent fake gradient
        net = synthetic.fill_state_dict(detector.VotingModule(1, 32), seed=3)
        for i, p in enumerate(net.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        list(net.parameters())[1].grad = None if rank == 0 else list(net.parameters())[1].grad
        nbytes = D.allreduce_gradients(net)
        grads = [p.grad.clone() for p in net.parameters()]
        mx = D.max_over_ranks(10.0 + rank, torch.device("cpu"))
        q.put((rank, nbytes, [float(g.flatten()[0]) for g in grads], mx, D.shard_scenes(33, rank, world)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, g0, m0, s0), (r1, n1, g1, m1, s1) = out
    assert g0 == g1                                   # identical averaged gradients on both ranks
    nparams = len(g0)
    for i in range(nparams):
        want = (1 * (i + 1) + 2 * (i + 1)) / 2.0
        if i == 1:
            want = (0 + 2 * 2) / 2.0                  # rank 0 had no gradient for param 1
        assert abs(g0[i] - want) < 1e-6, (i, g0[i], want)
    assert n0 == n1 and n0 > 0
    assert m0 == m1 == 11.0
    assert s0 == (0, 17) and s1 == (17, 33)           # contiguous, sizes differ by <= 1


def test_shard_scenes_partitions():
    from bridgeqa_b200.distributed import shard_scenes
    for total in (0, 1, 16, 17, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_scenes(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker_overlap(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bridgeqa_b200 import detector, distributed as D, synthetic
    try:
        torch.manual_seed(0)
        net = synthetic.fill_state_dict(detector.VotingModule(1, 32), seed=3)
        ref = synthetic.fill_state_dict(detector.VotingModule(1, 32), seed=3)
        reducer = D.OverlappedGradReducer(net, nbuckets=3)
        xyz = torch.randn(2, 16, 3, generator=torch.Generator().manual_seed(10 + rank))
        feat = torch.randn(2, 32, 16, generator=torch.Generator().manual_seed(20 + rank))
        launched_during_backward = []
        for step in range(2):                                  # the second step reuses the buckets
            reducer.prepare()
            v, f = net(xyz, feat)
            (v.sum() + f.pow(2).sum()).backward()
            launched_during_backward.append(sum(w is not None for w in reducer._works))
            nbytes = reducer.finish()
        # reference: local gradients, then the flat-bucket all-reduce of the same module
        v, f = ref(xyz, feat)
        (v.sum() + f.pow(2).sum()).backward()
        D.allreduce_gradients(ref)
        err = max(float((a.grad - b.grad).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
        q.put((rank, nbytes, err, launched_during_backward, len(reducer.buckets)))
    finally:
        dist.destroy_process_group()


def test_two_rank_overlapped_gradient_reducer_matches_flat_allreduce():
    """OverlappedGradReducer: buckets are launched from the backward hooks (all of them before finish()),
    the result is the mean the blocking flat-bucket all-reduce gives, and the buckets are reusable."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nbytes, err, launched, nb in out:
        assert nbytes > 0 and nb >= 2
        assert err < 1e-5, err
        assert launched == [nb, nb], launched            # every bucket went out during backward
