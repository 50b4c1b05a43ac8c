"""Per-kernel CUDA-event timing of the C-ABI calls, on the stream they are launched on.

    with KernelTimer() as kt:
        run_step()
    kt.summary()   # {"bqa_ball_query[n=40000,m=2048]": {"calls": 1, "ms": 0.31}, ...}

Events are recorded on torch's current stream immediately before and after each ABI call
(every ABI call enqueues exactly the kernels of one operator), so the measured interval
is that operator's device time inside the real step, not a cold isolated launch.
"""
import collections

import torch

from . import _native as N


class KernelTimer(object):
    def __init__(self, label_fn=None):
        self.records = []
        self._orig = None
        self._label_fn = label_fn

    def __enter__(self):
        self._orig = N.call
        timer = self

        def timed_call(name, *args):
            start = torch.cuda.Event(enable_timing=True)
            end = torch.cuda.Event(enable_timing=True)
            start.record()
            timer._orig(name, *args)
            end.record()
            ints = [a for a in args if isinstance(a, int)]
            timer.records.append((name, tuple(ints), start, end))

        N.call = timed_call
        return self

    def __exit__(self, *exc):
        N.call = self._orig
        return False

    def summary(self):
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for name, ints, start, end in self.records:
            key = "%s%s" % (name, list(ints))
            d = agg.setdefault(key, {"name": name, "dims": list(ints), "calls": 0, "ms": 0.0})
            d["calls"] += 1
            d["ms"] += start.elapsed_time(end)
        return agg
