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


def max_over_ranks(value, device, group=None):
    """max of a python float over the ranks (device-side timings are reported as the max)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
