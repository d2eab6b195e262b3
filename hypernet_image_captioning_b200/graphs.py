"""CUDA-graph capture of launch-bound inference loops (greedy decode: ~10 small launches per time step).

``graphed_call(key, fn, inputs)`` runs ``fn(*inputs)`` once eagerly per key to warm up, captures it into a CUDA graph on
static copies of ``inputs`` and afterwards only copies the inputs in and replays.  Outputs are cloned so callers own them
(the reference returns fresh tensors).  Anything that cannot be captured falls back to eager execution for that key.
"""
import warnings
from typing import Callable, Dict, Hashable, List, Sequence

import torch

ENABLED = True
_cache: Dict[Hashable, object] = {}


def clear():
    _cache.clear()


def graphed_call(key: Hashable, fn: Callable, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    if not ENABLED:
        out = fn(*inputs)
        return list(out) if isinstance(out, (tuple, list)) else [out]
    entry = _cache.get(key)
    if entry is None:
        static_in = [t.detach().clone() for t in inputs]
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn(*static_in)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn(*static_in)
            out = list(out) if isinstance(out, (tuple, list)) else [out]
            entry = (g, static_in, out)
        except Exception as e:  # noqa: BLE001  -- capture is an optimisation only
            warnings.warn(f"CUDA graph capture failed for {key!r} ({type(e).__name__}: {e}); running eagerly")
            torch.cuda.synchronize()
            entry = "eager"
        _cache[key] = entry
    if entry == "eager":
        out = fn(*inputs)
        return list(out) if isinstance(out, (tuple, list)) else [out]
    g, static_in, out = entry
    for s, t in zip(static_in, inputs):
        s.copy_(t, non_blocking=True)
    g.replay()
    return [o.clone() for o in out]


class GraphedStep:
    """Whole training step (forward + backward [+ gradient all-reduce]) as ONE CUDA graph.

    A caption training step is ~100 launches, most of them 2-30 us kernels around the two multi-millisecond weight
    streams; enqueued from Python the host needs about as long as the device, so any host-side work per step (input
    copies, reading the loss) shows up as device idle time.  Captured once and replayed, the step costs the host one
    ``cudaGraphLaunch``.

        gstep = GraphedStep(step_fn, (pooled, caps, h0), params=model.parameters())
        loss = gstep(pooled, caps, h0)     # inputs copied device-to-device into the static buffers, graph replayed

    ``step_fn(*inputs)`` must do everything on the device (no ``.item()``, no pageable host copies) and may call
    ``zero_grad(set_to_none=True)``: the gradients then live in the graph's memory pool, are rewritten by every replay,
    and are re-attached to ``param.grad`` after each call.  The returned tensors are the graph's static outputs (valid
    until the next call).  Drop every reference to losses / outputs of earlier eager steps before constructing it: a live
    autograd graph keeps its AccumulateGrad nodes bound to the stream they were created on, which invalidates the
    capture.  If capture fails the step runs eagerly -- same kernels, launched one by one.
    """

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], params=None, warmup: int = 3,
                 release: Callable = None):
        """``release``: called before the warm-up and again before the capture to drop whatever keeps the previous
        step's autograd graph alive (``HyperNet.release_graph``: the captioner holds the generated weights, and through
        them the hypernet's graph, until the next ``forward``)."""
        from . import _cabi
        self.fn = fn
        self.params = list(params) if params is not None else []
        self.static_in = [t.detach().clone() for t in example_inputs]
        self.graph = None
        self.launches_per_step = None
        self.static_out = None
        self.grads = []
        if not ENABLED:
            return
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            if release is not None:
                release()
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    fn(*self.static_in)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            if release is not None:
                release()
            for p in self.params:
                p.grad = None
            l0 = _cabi.launches()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn(*self.static_in)
            self.launches_per_step = _cabi.launches() - l0          # kernels launched by the library inside the capture
            self.static_out = out
            self.grads = [(p, p.grad) for p in self.params if p.grad is not None]
            self.graph = g
        except Exception as e:  # noqa: BLE001  -- capture is an optimisation only
            warnings.warn(f"CUDA graph capture of the training step failed ({type(e).__name__}: {e}); running eagerly")
            torch.cuda.synchronize()
            self.graph = None

    @property
    def captured(self) -> bool:
        return self.graph is not None

    def __call__(self, *inputs):
        if self.graph is None:
            return self.fn(*inputs)
        for s, t in zip(self.static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        for p, gr in self.grads:
            if p.grad is not gr:
                p.grad = gr
        return self.static_out
