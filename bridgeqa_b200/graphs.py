"""CUDA-graph replay of an inference forward (fixed input shape).

One detector forward is ~35 kernel launches on three streams (main, lower-level sampling,
ball-query grids); issued from Python through ctypes that is 2-4 ms of host time for a 2.5 ms
step.  Capturing the forward once and replaying it makes the host cost one cudaGraphLaunch.

    net = Pointnet2Backbone(...).cuda().eval()
    net.enable_cuda_graph()
    out = net({"point_clouds": pc})       # first call per shape: 2 eager warm-ups + capture

The tensors in `out` are STATIC buffers owned by the graph: the next call with the same graph
overwrites them (copy what must outlive it), exactly like torch.cuda.make_graphed_callables.  A
graph is re-captured when the input shape, the device, the operand precision or any
parameter/buffer of the module changes; training mode, autograd and CPU tensors go through the
eager path.

enable_cuda_graph(bind_inputs=True): no staging copy -- each distinct input BUFFER (data_ptr)
gets its own graph that reads the caller's tensor in place (a loader that cycles through a few
fixed device buffers, e.g. double-buffered H2D copies).  The graph keeps the tensor alive.
"""
import torch

from . import fused


class GraphedForward(object):
    MAX_GRAPHS = 32

    def __init__(self, module, impl, bind_inputs=False):
        self.module = module
        self.impl = impl            # callable(data_dict) -> data_dict, the eager forward
        self.bind_inputs = bind_inputs
        self.cache = {}
        self.replays = 0

    def applicable(self, data_dict):
        pc = data_dict.get("point_clouds")
        return ((torch.is_tensor(pc) or getattr(pc, "is_staged", False)) and pc.is_cuda
                and pc.dtype == torch.float32 and pc.dim() == 3
                and not self.module.training and not torch.is_grad_enabled()
                and fused.enabled() and set(data_dict.keys()) == {"point_clouds"})

    def _signature(self):
        return (fused.precision(),) + tuple(
            (t.data_ptr(), t._version)
            for t in list(self.module.parameters()) + list(self.module.buffers()))

    def __call__(self, data_dict):
        pc = data_dict["point_clouds"]
        bind = self.bind_inputs and pc.is_contiguous()
        # the sampling variant (latency / throughput, fused.lean_sampling) is baked into a graph
        key = (tuple(pc.shape), pc.device.index, pc.data_ptr() if bind else None, fused.lean_sampling_enabled())
        sig = self._signature()
        ent = self.cache.get(key)
        if ent is None or ent["sig"] != sig:
            if len(self.cache) >= self.MAX_GRAPHS:
                # the evicted graph (and the pool its buffers live in) may still be replaying
                torch.cuda.synchronize(pc.device)
                self.cache.pop(next(iter(self.cache)))
            ent = self._capture(pc, sig, bind)
            self.cache[key] = ent
        if pc.data_ptr() != ent["inp"].data_ptr():
            ent["inp"].copy_(pc if pc.is_contiguous() else pc.contiguous())
        ent["graph"].replay()
        self.replays += 1
        data_dict.update(ent["out"])
        return data_dict

    def _capture(self, pc, sig, bind):
        inp = pc.detach() if bind else pc.detach().clone(memory_format=torch.contiguous_format)
        with torch.cuda.device(pc.device):
            # eager warm-up on a side stream: lazily built state (folded weights, launch
            # attributes, occupancy plans) must exist before capture
            warm = torch.cuda.Stream(pc.device)
            warm.wait_stream(torch.cuda.current_stream(pc.device))
            with torch.cuda.stream(warm):
                for _ in range(2):
                    self.impl({"point_clouds": inp})
            torch.cuda.current_stream(pc.device).wait_stream(warm)
            torch.cuda.synchronize(pc.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = self.impl({"point_clouds": inp})
        out = {k: v for k, v in out.items() if k != "point_clouds"}
        return {"graph": graph, "inp": inp, "out": out, "sig": sig}


class Ticket(object):
    """One submitted forward: `out` (the module's data_dict) is valid once `done` has fired."""

    def __init__(self, out, done, stream, static):
        self.out, self.done, self.stream, self.static = out, done, stream, static

    def wait(self, stream=None):
        """Make `stream` (default: the current one) wait for this forward; returns the data_dict."""
        dev = self.stream.device
        stream = torch.cuda.current_stream(dev) if stream is None else stream
        stream.wait_event(self.done)
        if not self.static:
            # eager outputs belong to the caching allocator on the forward's stream
            for t in self.out.values():
                if torch.is_tensor(t) and t.is_cuda:
                    t.record_stream(stream)
        return self.out


class InFlight(object):
    """Keeps `depth` inference forwards of a module in flight, each on its own stream.

    The sampling chain of a 40k-point batch holds 96 of the 148 SMs for ~60 % of its forward and
    is pure SM-to-SM latency (2047 dependent argmax steps); the SA/FP kernels that follow are
    short.  Consecutive batches are independent, so the next batch's chain is issued on another
    stream and runs underneath the current batch's SA/FP kernels (measured on B200, graphed
    backbone, 16 x 40k: 2.03 ms/step serial, 1.53 with two batches in flight, 1.46 with three;
    outputs bit-identical).  In that regime SM-time, not the latency of one chain, bounds the
    throughput, so by default (lean_sampling=None -> depth > 1) the forwards run the throughput
    variant of the sorted sampling kernel (fused.lean_sampling: a 40k-point scene on 3 SMs instead
    of 6, same bits): 1.21 ms/step with three in flight.

        q = net.in_flight(depth=3)
        t = q.submit({"point_clouds": pc})        # returns at once
        ...
        out = t.wait()                            # current stream now waits for that forward

    With enable_cuda_graph(bind_inputs=True) every input buffer owns a graph and its static
    outputs, so forwards of different buffers overlap freely; a buffer (and its outputs) must not
    be reused before the ticket of its previous forward has been waited for."""

    def __init__(self, module, depth=2, lean_sampling=None):
        if depth < 1:
            raise ValueError("in_flight depth must be >= 1")
        self.module, self.depth = module, int(depth)
        self.lean = depth > 1 if lean_sampling is None else bool(lean_sampling)
        self._streams = {}
        self.submitted = 0

    def streams(self, device):
        key = torch.device(device).index
        if key not in self._streams:
            self._streams[key] = [torch.cuda.Stream(device) for _ in range(self.depth)]
        return self._streams[key]

    def submit(self, data_dict, after=None):
        """Issue module(data_dict) on the next stream of the ring.  `after`: event or list of
        events the forward has to wait for (default: everything queued on the current stream)."""
        pc = data_dict["point_clouds"]
        if not ((torch.is_tensor(pc) or getattr(pc, "is_staged", False)) and pc.is_cuda):
            raise RuntimeError("in-flight forwards need a CUDA point cloud (no CPU path)")
        s = self.streams(pc.device)[self.submitted % self.depth]
        self.submitted += 1
        if after is None:
            after = torch.cuda.Event()
            after.record(torch.cuda.current_stream(pc.device))
        for ev in (after if isinstance(after, (list, tuple)) else (after,)):
            s.wait_event(ev)
        g = getattr(self.module, "_graphed", None)
        if (g is not None and self.depth > 1 and g.applicable(data_dict)
                and not (g.bind_inputs and pc.is_contiguous())):
            # one shared staging buffer + one set of static outputs per shape: a second forward
            # in flight would overwrite both while the first still runs
            raise RuntimeError(
                "in_flight(depth > 1) over a CUDA graph needs enable_cuda_graph(bind_inputs=True) and "
                "contiguous inputs: with a shared staging buffer the forwards in flight would overwrite "
                "each other's input and static outputs")
        with torch.cuda.stream(s), torch.no_grad(), fused.lean_sampling(self.lean):
            static = g is not None and g.applicable(data_dict)
            if not static:
                pc.record_stream(s)
            out = self.module(data_dict)
            done = torch.cuda.Event()
            done.record(s)
        return Ticket(out, done, s, static)

    def drain(self, stream=None):
        """`stream` (default: current) waits for every forward submitted so far."""
        for dev, ring in self._streams.items():
            cur = torch.cuda.current_stream(ring[0].device) if stream is None else stream
            for s in ring:
                cur.wait_stream(s)
