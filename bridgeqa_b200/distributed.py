"""Multi-GPU plumbing of the hot path.

Scenes are independent units: the forward shards the batch dimension across ranks and uses
NO collective.  Training adds exactly one exchange: the all-reduce (mean) of the detector's
~0.95 M fp32 gradients (3.8 MB), which the reference gets from torch DDP
(/root/reference/scripts/train.py:347).  Here it is one flat bucket, so the NCCL call is a
single latency-bound launch over NVLink instead of one per parameter.  BatchNorm statistics
stay rank-local, like the reference (no SyncBN).
"""
import torch
import torch.distributed as dist


def shard_scenes(num_scenes, rank, world_size):
    """Contiguous [begin, end) slice of the batch for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(num_scenes), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_gradients(module, group=None):
    """Average the gradients of `module` over the process group through ONE flat buffer.
    Parameters without a gradient contribute zeros (DDP's find_unused_parameters=True
    behaviour, scripts/train.py:347).  Returns the number of bytes reduced."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float()
                      for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    offset = 0
    for p in params:
        n = p.numel()
        g = flat[offset:offset + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        offset += n
    return flat.numel() * 4


class OverlappedGradReducer(object):
    """Gradient all-reduce launched from inside the backward pass ("overlappable
    with the SA1 backward"; the reference gets the same from DDP's bucket hooks, scripts/train.py:347).

    The parameters are split, in reverse registration order (~ the order their gradients become final),
    into `nbuckets` flat fp32 buffers; every parameter's .grad is a view into its bucket, so autograd
    accumulates straight into the buffer that goes on the wire.  A post-accumulate hook counts the
    parameters of a bucket as they finish; the last one launches that bucket's NCCL all-reduce
    asynchronously, so the buckets of the late layers (proposal, voting, FP, SA4-2) travel while SA1's
    backward -- the largest part of the step -- is still running.  finish() launches whatever never got
    a gradient (DDP's find_unused_parameters behaviour), waits, and leaves the MEAN in every .grad.

        reducer = OverlappedGradReducer(model)          # once
        reducer.prepare(); loss.backward(); reducer.finish()
    """

    def __init__(self, module, group=None, nbuckets=2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        params = [p for p in module.parameters() if p.requires_grad]
        if any(p.dtype != torch.float32 for p in params):
            raise RuntimeError("OverlappedGradReducer: fp32 parameters only (the buckets are flat fp32 buffers "
                               "that the gradients are views of)")
        self.params = list(reversed(params))
        total = sum(p.numel() for p in self.params)
        target = max(1, -(-total // max(1, int(nbuckets))))
        self.buckets = []                               # (flat, [(param, offset)])
        cur, cur_n = [], 0
        for p in self.params:
            cur.append(p)
            cur_n += p.numel()
            if cur_n >= target:
                self._close(cur)
                cur, cur_n = [], 0
        if cur:
            self._close(cur)
        self._bucket_of = {}
        for bi, (_flat, members) in enumerate(self.buckets):
            for p, _off in members:
                self._bucket_of[id(p)] = bi
        self._pending = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._armed = False
        self.defer = False          # True: hooks only count; finish() launches (CUDA-graph capture of the step)
        self.bytes = 4 * total
        self._avg = self.world > 1 and dist.get_backend(group) == "nccl"
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    def _close(self, members):
        dev = members[0].device
        flat = torch.zeros(sum(p.numel() for p in members), dtype=torch.float32, device=dev)
        off, table = 0, []
        for p in members:
            table.append((p, off))
            off += p.numel()
        self.buckets.append((flat, table))

    def prepare(self):
        """zero the buckets and point every .grad into them (call before backward)"""
        for bi, (flat, members) in enumerate(self.buckets):
            flat.zero_()
            for p, off in members:
                p.grad = flat[off:off + p.numel()].view_as(p)
            self._pending[bi] = len(members)
            self._works[bi] = None
        self._armed = True

    def _launch(self, bi):
        if self.world > 1 and self._works[bi] is None:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._works[bi] = dist.all_reduce(self.buckets[bi][0], op=op, group=self.group, async_op=True)

    def _hook(self, p):
        if not self._armed:
            return
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and not self.defer:
            self._launch(bi)

    def attach(self):
        """point every .grad into the buckets without zeroing them (after a graph replay wrote them)"""
        for flat, members in self.buckets:
            for p, off in members:
                p.grad = flat[off:off + p.numel()].view_as(p)
        self._works = [None] * len(self.buckets)

    def finish(self):
        """-> bytes reduced; every .grad then holds the mean over the ranks"""
        self._armed = False
        if self.world <= 1:
            return 0
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for bi, w in enumerate(self._works):
            w.wait()
            if not self._avg:
                self.buckets[bi][0].div_(self.world)
        return self.bytes


def max_over_ranks(value, device, group=None):
    """max of a python float over the ranks (device-side timings are reported as the max)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
