"""DET-stage training step of the detector front-end (BASELINE.json configs[3]): forward in
train mode (batch-statistics BN, un-fused differentiable operators), backward through the
four gradient kernels (gather / group / three_interpolate scatter-adds) and the torch MLPs,
then ONE flat-bucket NCCL all-reduce of the gradients (bridgeqa_b200.distributed).

The dataset losses of the reference (lib/loss_helper.py) are out of scope; the step uses the
stand-in SURVEY.md section 8d prescribes: a fixed random projection of `fp2_features` and
`aggregated_vote_features`, so every parameter on the path receives a gradient.
"""
import torch
import torch.distributed as dist

from . import distributed as D


class ProjectionLoss(torch.nn.Module):
    def __init__(self, seed=0):
        super().__init__()
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        self.register_buffer("p_seed", torch.randn(256, generator=gen) / 16.0)
        self.register_buffer("p_vote", torch.randn(128, generator=gen) / 16.0)

    def forward(self, data_dict):
        a = torch.einsum("bcn,c->bn", data_dict["fp2_features"], self.p_seed)
        b = torch.einsum("bkc,c->bk", data_dict["aggregated_vote_features"], self.p_vote)
        c = data_dict["center"].sum(-1) * 1e-3 + data_dict["objectness_scores"].sum(-1) * 1e-3
        return a.pow(2).mean() + b.pow(2).mean() + c.pow(2).mean()


def train_step(model, loss_fn, point_clouds, optimizer=None, next_point_clouds=None, reducer=None):
    """One fwd + bwd (+ gradient all-reduce when a process group is up, + optimizer step).
    Returns the detached loss.

    next_point_clouds: the batch the NEXT call will be given (already on the device).  Its SA1
    sampling -- 1.3 ms of serial chain that depends on coordinates only, not on the weights this
    step updates -- is issued on a side stream before this step's backward and runs underneath it
    (Pointnet2Backbone.prefetch_sampling); the next forward picks it up.

    reducer: a distributed.OverlappedGradReducer built once for `model` -- the gradient all-reduce is then
    launched bucket by bucket from inside the backward pass instead of after it."""
    model.train()
    if reducer is not None:
        reducer.prepare()
    elif optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    else:
        for p in model.parameters():
            p.grad = None
    out = model({"point_clouds": point_clouds})
    loss = loss_fn(out)
    if next_point_clouds is not None:
        backbone = getattr(model, "detection_backbone", model)
        if hasattr(backbone, "prefetch_sampling"):
            backbone.prefetch_sampling(next_point_clouds)
    loss.backward()
    if reducer is not None:
        reducer.finish()
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        D.allreduce_gradients(model)
    if optimizer is not None:
        torch.nn.utils.clip_grad_value_(model.parameters(), 1.0)     # lib/solver.py:409
        optimizer.step()
    return loss.detach()
