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
    MAX_GRAPHS = 12

    def __init__(self, module, impl, bind_inputs=False):
        self.module = module
        self.impl = impl            # callable(data_dict) -> data_dict, the eager forward
        self.bind_inputs = bind_inputs
        self.cache = {}
        self.replays = 0

    def applicable(self, data_dict):
        pc = data_dict.get("point_clouds")
        return (torch.is_tensor(pc) and pc.is_cuda and pc.dtype == torch.float32 and pc.dim() == 3
                and not self.module.training and not torch.is_grad_enabled()
                and fused.enabled() and set(data_dict.keys()) == {"point_clouds"})

    def _signature(self):
        return (fused.precision(),) + tuple(
            (t.data_ptr(), t._version)
            for t in list(self.module.parameters()) + list(self.module.buffers()))

    def __call__(self, data_dict):
        pc = data_dict["point_clouds"]
        bind = self.bind_inputs and pc.is_contiguous()
        key = (tuple(pc.shape), pc.device.index, pc.data_ptr() if bind else None)
        sig = self._signature()
        ent = self.cache.get(key)
        if ent is None or ent["sig"] != sig:
            if len(self.cache) >= self.MAX_GRAPHS:
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
