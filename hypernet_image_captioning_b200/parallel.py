"""Batch data-parallel plumbing (one process per GPU, torch.distributed / NCCL over NVLink).

The path shards by batch only: captions are independent given (theta, shared parameters).  The
hypernetwork and its inputs are replicated, so instead of all-reducing the (multi-GB) head-parameter gradients the
exchange happens at the bottleneck: the gradient w.r.t. the generated weights, d(theta) [G, theta] (0.6-1.4 MB per
style group).  Every rank then runs the same head backward on the same summed d(theta) and obtains identical head
gradients with no further communication.  The shared decoder parameters (embedding, vocabulary projection, feature
layers) are all-reduced in one flat bucket after the backward pass.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch.autograd import Function


class AllReduceGradFn(Function):
    """Identity in the forward pass; sums the gradient over the process group in the backward pass."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)   # the head backward waits for this one: issue it first
        flush_ready()      # the decoder's shared gradients are final by now: reduce them on the comm stream meanwhile
        return g, None


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def allreduce_grad(x: torch.Tensor, group=None) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return AllReduceGradFn.apply(x, group)


def shared_parameters(model) -> List[torch.nn.Parameter]:
    """Parameters whose gradients differ across ranks: everything except the hypernetwork (hn_base / hn_heads)."""
    return [p for n, p in model.named_parameters() if not (n.startswith("hn_base.") or n.startswith("hn_heads."))]


class ScaleGradFn(Function):
    """Identity forward; multiplies the gradient by a constant.  Used on the hypernet INPUT under data parallelism: the
    style vector's gradient comes out of the (replicated) hypernet backward already global, so it must enter the summed
    shared-gradient bucket as 1/world of itself (the style is a row of ``captioner.embed.weight`` in the Flickr setting,
    hypernet_attention.py:139-142)."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


def _flat_allreduce(grads, group, buf=None):
    """SUM all-reduce of a list of tensors through ONE flat buffer (copy in, all-reduce, copy back)."""
    n = sum(g.numel() for g in grads)
    if buf is None or buf.numel() < n or buf.device != grads[0].device or buf.dtype != grads[0].dtype:
        buf = torch.empty(n, device=grads[0].device, dtype=grads[0].dtype)
    flat = buf[:n]
    off = 0
    for g in grads:
        flat[off:off + g.numel()].copy_(g.reshape(-1))
        off += g.numel()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return buf


class _Overlap:
    """State of the overlapped shared-gradient all-reduce (``enable_overlap``).

    Round 1 all-reduced the 15 MB shared bucket after the whole backward (SCALE_r01: 95.6 % at 8 GPUs).  Those gradients
    are complete as soon as the decoder's backward node has run, i.e. BEFORE the hypernet head backward that dominates
    the step (2.1 of 4.5 ms).  Post-accumulate-grad hooks collect the parameters whose ``.grad`` is final; the hypernet's
    backward calls ``flush_ready()`` as its first action, which packs them into a persistent flat buffer and all-reduces
    it on a side ("comm") stream and a dedicated communicator (so it does not queue behind -- or in front of -- the
    d(theta) all-reduce the head backward is waiting for), then ``join()`` as its last action, so everything that runs
    after the hypernet backward (late gradient contributions, the optimizer) is ordered after the reduction.
    ``allreduce_shared_grads`` joins and reduces whatever became ready later (the embedding matrix when the style vector
    is one of its rows) on the current stream."""

    def __init__(self):
        self.enabled = False
        self.params = []
        self.ids = set()
        self.ready = []
        self.done = set()
        self.group = None
        self.stream = None
        self.buf = None
        self.tail_buf = None
        self.inflight = False
        self.handles = []
        self.src_streams = []
        self.routes = []


_ov = _Overlap()
_own_group = None     # the dedicated communicator survives enable/disable cycles (creating one is a collective + ~100 ms)


def enable_overlap(params: Iterable[torch.nn.Parameter], group=None, own_communicator: bool = True):
    """Overlap the shared-gradient all-reduce of ``params`` with the hypernet head backward (see ``_Overlap``).
    Collective: every rank must call it (it may create a process group).  ``disable_overlap()`` undoes it."""
    disable_overlap()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return False
    _ov.params = list(params)
    _ov.ids = {id(p) for p in _ov.params}
    _ov.group = group
    if own_communicator and group is None:
        global _own_group
        if _own_group is None:
            _own_group = dist.new_group(list(range(dist.get_world_size())))
        _ov.group = _own_group
    cuda = any(p.is_cuda for p in _ov.params)
    _ov.stream = torch.cuda.Stream() if cuda else None
    for p in _ov.params:
        _ov.handles.append(p.register_post_accumulate_grad_hook(_on_grad_ready))
    _ov.enabled = True
    return True


def disable_overlap():
    for h in _ov.handles:
        h.remove()
    _ov.__init__()


def route_style_grad(x: torch.Tensor):
    """Overlap mode only.  When the hypernet input ``x`` is a contiguous view of a shared leaf parameter (the Flickr idiom:
    the style vector is a row of ``captioner.embed.weight``, hypernet_attention.py:139-142), that parameter would receive a
    second gradient contribution at the very end of the backward (through the hypernet) and could only be all-reduced after
    it.  Instead the hypernet runs on a detached leaf copy of ``x``; the parameter's gradient is then final as soon as the
    decoder's backward has run (so it joins the overlapped bucket), and ``allreduce_shared_grads`` adds the leaf's gradient
    -- identical on every rank, hence added once, unscaled -- into the reduced gradient.  Returns the leaf, or None."""
    base = getattr(x, "_base", None)
    if not _ov.enabled or base is None or not (base.is_leaf and base.requires_grad) or id(base) not in _ov.ids:
        return None
    if not (x.is_contiguous() and base.is_contiguous() and x.dtype == base.dtype):
        return None
    leaf = x.detach().requires_grad_(True)
    _ov.routes.append((base, x.storage_offset(), leaf))
    return leaf


def _apply_routes():
    for base, off, leaf in _ov.routes:
        if leaf.grad is not None and base.grad is not None:
            base.grad.view(-1)[off:off + leaf.numel()].add_(leaf.grad.reshape(-1))
    _ov.routes = []


def _on_grad_ready(p):
    if _ov.enabled and id(p) not in _ov.done:
        _ov.ready.append(p)
        if p.is_cuda:       # the gradient was accumulated on this stream (feature-branch parameters: a side stream)
            s = torch.cuda.current_stream()
            if s not in _ov.src_streams:
                _ov.src_streams.append(s)


def flush_ready():
    """All-reduce (on the comm stream) every shared gradient that is final so far.  Called by the hypernet backward."""
    if not _ov.enabled or not _ov.ready:
        return
    ps = [p for p in _ov.ready if p.grad is not None]
    _ov.ready = []
    if not ps:
        return
    for p in ps:
        _ov.done.add(id(p))
    grads = [p.grad for p in ps]
    if _ov.stream is None:
        _ov.buf = _flat_allreduce(grads, _ov.group, _ov.buf)
        return
    _ov.stream.wait_stream(torch.cuda.current_stream())
    for src in _ov.src_streams:
        _ov.stream.wait_stream(src)
    _ov.src_streams = []
    with torch.cuda.stream(_ov.stream):
        _ov.buf = _flat_allreduce(grads, _ov.group, _ov.buf)
    _ov.inflight = True


def join():
    """Order the current stream after the in-flight bucket (no host synchronisation)."""
    if _ov.enabled and _ov.inflight and _ov.stream is not None:
        torch.cuda.current_stream().wait_stream(_ov.stream)
        _ov.inflight = False


def allreduce_shared_grads(params: Iterable[torch.nn.Parameter], group=None, extra: Optional[torch.Tensor] = None):
    """SUM all-reduce of the gradients of ``params`` (+ an optional extra tensor, e.g. the token count) -- call after
    ``loss.backward()``.  With ``enable_overlap`` most of them were already reduced during the backward; this joins that
    reduction and handles the rest in one flat bucket on the current stream."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    params = list(params)
    if _ov.enabled:
        join()
        # whatever became ready after the hypernet backward started (or when no hypernet backward ran) was never
        # flushed: it is reduced here together with the parameters that are not in the overlapped set at all
        rest = [p for p in params if id(p) not in _ov.done]
        _ov.ready, _ov.done, _ov.src_streams = [], set(), []
        group = _ov.group if group is None else group
    else:
        rest = params
    grads = [p.grad for p in rest if p.grad is not None]
    if extra is not None:
        grads = grads + [extra]
    if grads:
        _ov.tail_buf = _flat_allreduce(grads, group, _ov.tail_buf)
    if _ov.enabled:
        _apply_routes()


def loss_weight(captions: torch.Tensor, ignore_index: Optional[int] = None, group=None) -> torch.Tensor:
    """Weight that turns this rank's mean caption loss into its share of the GLOBAL-batch mean:
    ``n_valid(local) / n_valid(all ranks)`` (one 8-byte all-reduce).  With ``ignore_index`` (cc_train_hypernet.py:153)
    ranks hold different numbers of non-pad tokens, so the plain 1/N average of per-rank means is not the global mean;
    ``(loss * loss_weight(caps, pad)).backward()`` on every rank followed by the gradient all-reduces reproduces the
    single-process gradients of the concatenated batch exactly.  Without ``ignore_index`` it is B_local / B_global."""
    n = (captions != ignore_index).sum() if ignore_index is not None else torch.tensor(captions.numel(), device=captions.device)
    n = n.to(torch.float64).reshape(1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return torch.ones((), device=captions.device)
    tot = n.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return (n / tot).to(torch.float32).reshape(())
