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


class GraphedTrainStep(object):
    """fwd + bwd of the DET training step captured ONCE into two alternating CUDA graphs and replayed.

    Issued eagerly the step is ~1000 kernel launches: 10.4 ms of kernels take 13.9 ms on an idle host and
    15-50 ms on a loaded one.  What makes whole-step capture non-trivial here is the sampling prefetch
    (SA1's furthest point sampling of batch k+1 runs under the backward of batch k): a replay has to
    PRODUCE the sampling the next replay CONSUMES.  So there are two static input buffers and two graphs:

        graph 0:  forward + backward on buffer 0 using sampling record 0  ||  sampling of buffer 1 -> record 1
        graph 1:  forward + backward on buffer 1 using sampling record 1  ||  sampling of buffer 0 -> record 0

    Gradients land in the static flat buckets of a distributed.OverlappedGradReducer (every .grad is a view
    into them), which are all-reduced after the replay when a process group is up.  The optimizer step stays
    outside the graph.  BatchNorm momentum values are baked into the capture: a change (BNMomentumScheduler,
    lib/solver.py:271-279) triggers a re-capture.

        step = GraphedTrainStep(model, loss_fn, example_batch, optimizer)
        for pc, pc_next in batches:           # pc_next: the batch the NEXT call gets (or None)
            loss = step(pc, pc_next)
    """

    def __init__(self, model, loss_fn, example, optimizer=None, reducer=None):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.backbone = getattr(model, "detection_backbone", model)
        self.reducer = reducer if reducer is not None else D.OverlappedGradReducer(model)
        self.reducer.defer = True
        self.bufs = [torch.empty_like(example), torch.empty_like(example)]
        self.rec = [None, None]          # static sampling records
        self.graphs = [None, None]
        self.losses = [None, None]
        self.k = 0
        self._staged = None              # (tensor, version) whose data sits in bufs[k] with a valid record
        self._sig = None
        self.bufs[0].copy_(example)
        self.bufs[1].copy_(example)
        model.train()
        # eager warm-up on a side stream (library handles, autotuning, lazily created streams / maps); the
        # BatchNorm running statistics it moves are put back afterwards
        saved = [b.detach().clone() for b in model.buffers()]
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager(self.bufs[0])
        torch.cuda.current_stream(example.device).wait_stream(side)
        with torch.no_grad():
            for b, old in zip(model.buffers(), saved):
                b.copy_(old)
        for i in (0, 1):
            self.rec[i] = self._sample(self.bufs[i])
        torch.cuda.synchronize(example.device)

    # -- pieces --------------------------------------------------------------------------------
    def _eager(self, pc):
        self.reducer.prepare()
        loss = self.loss_fn(self.model({"point_clouds": pc}))
        loss.backward()
        return loss

    def _sample(self, pc):
        """SA1's sampling record of `pc`, computed now (fresh tensors)."""
        if not self.backbone.prefetch_sampling(pc):
            return None
        pre, self.backbone._prefetched = self.backbone._prefetched, None
        torch.cuda.current_stream(pc.device).wait_event(pre["done"])
        return pre

    def _signature(self):
        return tuple(float(m.momentum) for m in self.model.modules()
                     if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.momentum is not None)

    def _capture(self):
        dev = self.bufs[0].device
        self._sig = self._signature()
        pool = None
        from . import _native
        for k in (0, 1):
            before = _native.launch_count()
            g = torch.cuda.CUDAGraph()
            cur, nxt = self.bufs[k], self.bufs[1 - k]
            with torch.cuda.graph(g, pool=pool):
                main = torch.cuda.current_stream(dev)
                if self.rec[1 - k] is not None:
                    # the NEXT batch's sampling, on the side stream, into the other record's tensors
                    made = self._sample_captured(nxt, self.rec[1 - k])
                if self.rec[k] is not None:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    self.backbone._prefetched = dict(self.rec[k], pc=cur, version=cur._version, done=ev)
                loss = self._eager(cur)
                if self.rec[1 - k] is not None:
                    main.wait_event(made)
            self.graphs[k], self.losses[k] = g, loss
            self.launches_per_step = _native.launch_count() - before     # C-ABI launches one replay re-issues
            pool = g.pool()
        self.backbone._prefetched = None

    def _sample_captured(self, pc, rec):
        ok = self.backbone.prefetch_sampling(pc)
        assert ok
        pre, self.backbone._prefetched = self.backbone._prefetched, None
        side = fused_side_stream(pc.device)
        with torch.cuda.stream(side), torch.no_grad():
            rec["xyz"].copy_(pre["xyz"])
            rec["grid"][0].copy_(pre["grid"][0])
            rec["inds"].copy_(pre["inds"])
            rec["new_xyz"].copy_(pre["new_xyz"])
            made = torch.cuda.Event()
            made.record(side)
        return made

    # -- the step ------------------------------------------------------------------------------
    def __call__(self, point_clouds, next_point_clouds=None):
        self.model.train()
        if self.graphs[0] is None or self._sig != self._signature():
            self._capture()
        k = self.k
        st = self._staged
        if st is None or st[0] is not point_clouds or st[1] != point_clouds._version:
            # not announced by the previous call: stage and sample it now
            self.bufs[k].copy_(point_clouds)
            if self.rec[k] is not None:
                fresh = self._sample(self.bufs[k])
                for key in ("xyz", "inds", "new_xyz"):
                    self.rec[k][key].copy_(fresh[key])
                self.rec[k]["grid"][0].copy_(fresh["grid"][0])
        if next_point_clouds is not None:
            self.bufs[1 - k].copy_(next_point_clouds)
            self._staged = (next_point_clouds, next_point_clouds._version)
        else:
            self._staged = None
        self.graphs[k].replay()
        self.reducer.attach()
        self.reducer.finish()
        if self.optimizer is not None:
            torch.nn.utils.clip_grad_value_(self.model.parameters(), 1.0)     # lib/solver.py:409
            self.optimizer.step()
        self.k = 1 - k
        return self.losses[k].detach()


def fused_side_stream(device):
    from . import fused
    return fused.side_stream(device, "prefetch")
