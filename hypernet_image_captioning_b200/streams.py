"""Side streams for the independent branches of a step (capture-safe: plain fork / join with stream waits).

The attention-variant step has three branches that only meet at the recurrence:

    hypernet   x -> theta                       HBM-bound weight streaming        ("hypernet" stream, opt-in)
    features   feature_fc -> f, K = W_a f, h0   tensor-core GEMMs                 ("features" stream)
    recurrence + vocabulary projection + loss   everything else                   (the caller's stream)

Each branch is its own autograd node created under its stream, so the autograd engine runs the matching backward nodes
on the same streams (and inserts the cross-stream waits itself): the head backward (HBM) and the feature_fc backward
(tensor cores) overlap after the recurrence's BPTT.  Every use of a side stream starts with ``side.wait_stream(current)``,
so memory handed between streams is never reused early.  ``CAPHN_OVERLAP=0`` keeps everything on the caller's stream.
"""
import os
from typing import Dict, List, Tuple

import torch

# "capture" (default): side streams only while a CUDA graph is being captured -- replayed graphs get the parallel branches for
# free, while an eagerly launched step is bound by the host's launch rate and every fork / join adds host work (measured:
# the eager attention step went 147 k -> 127 k captions/s with the streams always on).  True / "always": also in eager mode
# (CAPHN_OVERLAP=always).  False: never (CAPHN_OVERLAP=0).
_env = os.environ.get("CAPHN_OVERLAP", "capture").lower()
ENABLED = False if _env in ("0", "off", "false") else (True if _env in ("always", "2") else "capture")
_streams: Dict[Tuple[str, int], "torch.cuda.Stream"] = {}
_pending: List["torch.cuda.Stream"] = []


def enabled(t: torch.Tensor = None) -> bool:
    if not ENABLED or (t is not None and not t.is_cuda):
        return False
    if ENABLED == "capture":
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    return True


def side(name: str, device=None, priority: int = 0) -> "torch.cuda.Stream":
    """Named side stream (one per device).  ``priority`` < 0: a high-priority stream -- its pending CTAs are scheduled before
    those of default-priority streams (used for the dependency chain of the decode loops)."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    key = (name, idx)
    if key not in _streams:
        _streams[key] = torch.cuda.Stream(device=idx, priority=priority)
    return _streams[key]


class fork:
    """``with fork("features") as s:`` -- run the block on side stream ``s`` after everything queued on the current stream.
    The caller joins with ``torch.cuda.current_stream().wait_stream(s)`` (or ``defer(s)`` + ``wait_pending()``)."""

    def __init__(self, name: str, priority: int = 0):
        self.name, self.priority = name, priority

    def __enter__(self):
        cur = torch.cuda.current_stream()
        self.s = side(self.name, priority=self.priority)
        self.s.wait_stream(cur)
        self.ctx = torch.cuda.stream(self.s)
        self.ctx.__enter__()
        return self.s

    def __exit__(self, *exc):
        return self.ctx.__exit__(*exc)


def defer(s: "torch.cuda.Stream"):
    """The current stream must wait for ``s`` before touching what the forked block produced; the wait is issued by the
    consumer (``wait_pending``), so work queued in between overlaps with the block."""
    if s not in _pending:
        _pending.append(s)


# Partial readiness of a tensor that several streams are still writing (the generated weight vector theta: every hypernet head
# writes its own slice on its own branch stream).  The producer marks each slice when it is complete; a consumer that needs
# only some slices waits for exactly those instead of for the whole forked computation (``wait_pending``).
_ranges: List[Tuple[int, int, "torch.cuda.Event"]] = []


def clear_ranges():
    _ranges.clear()


def mark_range(t: torch.Tensor):
    """The bytes of ``t`` (a contiguous view) are complete once everything queued so far on the current stream has run."""
    if not t.is_cuda:
        return
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    _ranges.append((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), ev))


def wait_ranges(tensors) -> bool:
    """Make the current stream wait for the marks that cover ``tensors``; False (nothing waited for) unless every tensor
    is fully covered by marked ranges."""
    todo = []
    for t in tensors:
        if not t.is_cuda:
            return False
        a, b = t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()
        hit = [r for r in _ranges if r[0] < b and a < r[1]]
        if not hit or min(r[0] for r in hit) > a or max(r[1] for r in hit) < b:
            return False
        todo += hit
    cur = torch.cuda.current_stream()
    for r in todo:
        cur.wait_event(r[2])
    return True


def wait_pending():
    if _pending:
        cur = torch.cuda.current_stream()
        for s in _pending:
            cur.wait_stream(s)
        _pending.clear()


class Branches:
    """Independent groups of small launches (weight-gradient GEMMs with their operand splits, bias sums, scatter-adds) on
    side streams, joined at exit.  The backward tails are chains of 5-30 us kernels whose grids do not fill the 148 SMs;
    run back to back they cost their summed latencies, side by side roughly the longest chain.

        with streams.Branches("tail", like=dGI) as br:
            with br.on(0): dW_ih = ops.matmul_tn(dGI, XC)
            with br.on(1): dW_hh = ops.matmul_tn(dGH, Hprev)
            demb = ...                                   # the caller's stream keeps working meanwhile

    ``on(i)``: the block runs on branch stream i, after everything queued SO FAR on the caller's stream (and after branch
    i's earlier blocks); the caller's stream waits for all branches at exit.  Disabled (``CAPHN_OVERLAP=0`` or a CPU
    tensor): the blocks simply run in order."""

    def __init__(self, name: str, like: torch.Tensor = None, enable: bool = True):
        self.name = name
        self.active = enable and enabled(like)
        self.used = {}

    def __enter__(self):
        if self.active:
            self.cur = torch.cuda.current_stream()
        return self

    def on(self, i: int):
        if not self.active:
            import contextlib
            return contextlib.nullcontext()
        s = self.used.get(i)
        if s is None:
            s = self.used[i] = side(f"{self.name}{i}")
        s.wait_stream(self.cur)
        return torch.cuda.stream(s)

    def join(self, i: int):
        """The caller's stream waits for what branch i has been given so far (a partial join before __exit__)."""
        if self.active and i in self.used:
            self.cur.wait_stream(self.used[i])

    def __exit__(self, *exc):
        if self.active:
            for s in self.used.values():
                self.cur.wait_stream(s)
        return False
