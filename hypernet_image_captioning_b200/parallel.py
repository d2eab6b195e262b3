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
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def allreduce_grad(x: torch.Tensor, group=None) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return AllReduceGradFn.apply(x, group)


def shared_parameters(model) -> List[torch.nn.Parameter]:
    """Parameters whose gradients differ across ranks: everything except the hypernetwork (hn_base / hn_heads)."""
    return [p for n, p in model.named_parameters() if not (n.startswith("hn_base.") or n.startswith("hn_heads."))]


def allreduce_shared_grads(params: Iterable[torch.nn.Parameter], group=None, extra: Optional[torch.Tensor] = None):
    """One flat SUM all-reduce over the gradients of ``params`` (+ an optional extra tensor, e.g. the token count)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if extra is not None:
        grads = grads + [extra]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


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
