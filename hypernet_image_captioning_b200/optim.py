"""Optimizer step of the reference trainers on the B200 kernels.

The reference builds `torch.optim.Adam(params, lr=self.hparams['lr'])` (cc_train_hypernet.py:110-122,
hypernet.py:116-123) and lets Lightning clip the global gradient norm (`gradient_clip_val=5.`,
cc_train_hypernet.py:405).  `FusedAdam` is a drop-in `torch.optim.Optimizer` with the same state (`step`, `exp_avg`,
`exp_avg_sq`), so `ReduceLROnPlateau` and `state_dict()` / `load_state_dict()` keep working; `max_grad_norm` folds the
clipping into the step: the norm is reduced on the device, the coefficient is applied to the gradient on the fly
(no host synchronisation, `.grad` is left untouched).

Rank-G head gradients: with `hypernet.head_grad_mode = "lowrank"` the backward leaves the gradient of the large head
matrices as the pair `(dtheta [G,N], a [G,K])` on `param.grad_lowrank` (G = number of style groups, 1 at every reference
call site) instead of writing the dense `dW = dtheta^T a`; `FusedAdam.step()` takes the norm from two G x G Gram matrices
and forms `g[n,k]` on the fly inside the update.  Per head parameter and training step: 4 (forward) + 4 (backward, dA only)
+ 24 (update) = 32 bytes of HBM traffic instead of 4 + 8 + 32 = 44.
"""
import torch

from . import _cabi
from .ops import _stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self._scratch = None          # [sumsq (double), coef (float), norm (float)] on the device
        self.last_grad_norm = None    # device scalar tensor (set when max_grad_norm is not None)

    def _scratch_on(self, dev):
        if self._scratch is None or self._scratch[0].device != dev:
            self._scratch = (torch.zeros(1, device=dev, dtype=torch.float64), torch.ones(1, device=dev),
                             torch.zeros(1, device=dev))
        return self._scratch

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none)
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "grad_lowrank", None) is not None:
                    p.grad_lowrank = None

    def _state_of(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
            if p.dtype == torch.bfloat16:      # bf16 mode: fp32 master weight, the bf16 parameter is its rounding
                st["master"] = p.detach().to(torch.float32).contiguous()
        st["step"] = int(st["step"]) + 1
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = []
        for group in self.param_groups:
            for p in group["params"]:
                lowrank = getattr(p, "grad_lowrank", None)
                if p.grad is None and lowrank is not None:
                    if p.dtype != torch.float32 or not p.is_contiguous() or p.dim() != 2:
                        raise _cabi.CaphnError("rank-G gradients need contiguous fp32 [N,K] parameters")
                    st = self._state_of(p)
                    items.append((group, p, lowrank, st))
                    continue
                if p.grad is None:
                    continue
                if p.dtype not in (torch.float32, torch.bfloat16) or not p.is_cuda or not p.is_contiguous():
                    raise _cabi.CaphnError("FusedAdam needs contiguous fp32 (or bf16, with fp32 master state) CUDA parameters")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.dtype != p.dtype:
                    g = g.to(p.dtype)
                st = self._state_of(p)
                items.append((group, p, g, st))
        if not items:
            return loss
        gscale = None
        if self.max_grad_norm is not None:
            sumsq, coef, norm = self._scratch_on(items[0][1].device)
            sumsq.zero_()
            for _, _, g, _ in items:
                if isinstance(g, tuple):      # ||dP^T a||_F^2 = sum_{q,r} (dP_q . dP_r)(a_q . a_r): no W-sized pass
                    dP, a = g
                    G = dP.shape[0]
                    gram = torch.zeros(2, G * G, device=dP.device, dtype=torch.float64)
                    _cabi.call("caphn_gram", dP.data_ptr(), dP.stride(0), G, dP.shape[1], gram[0].data_ptr(), _stream())
                    _cabi.call("caphn_gram", a.data_ptr(), a.stride(0), G, a.shape[1], gram[1].data_ptr(), _stream())
                    _cabi.call("caphn_sumsq_lowrank", gram[0].data_ptr(), gram[1].data_ptr(), G, sumsq.data_ptr(), _stream())
                elif g.dtype == torch.bfloat16:
                    _cabi.call("caphn_sumsq_bf16", g.data_ptr(), g.numel(), sumsq.data_ptr(), _stream())
                else:
                    _cabi.call("caphn_sumsq", g.data_ptr(), g.numel(), sumsq.data_ptr(), _stream())
            _cabi.call("caphn_clip_coef", sumsq.data_ptr(), float(self.max_grad_norm), coef.data_ptr(), norm.data_ptr(),
                       _stream())
            gscale, self.last_grad_norm = coef.data_ptr(), norm
        for group, p, g, st in items:
            b1, b2 = group["betas"]
            if isinstance(g, tuple):
                dP, a = g
                _cabi.call("caphn_adam_step_lowrank", p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                           dP.data_ptr(), dP.stride(0), a.data_ptr(), a.stride(0), dP.shape[0], p.shape[0], p.shape[1],
                           float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                           st["step"], gscale, _stream())
                p.grad_lowrank = None         # consumed (zero_grad() does not know about it)
                continue
            if p.dtype == torch.bfloat16:
                _cabi.call("caphn_adam_step_bf16", p.data_ptr(), g.data_ptr(), st["master"].data_ptr(),
                           st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]), float(b1),
                           float(b2), float(group["eps"]), float(group["weight_decay"]), st["step"], gscale, _stream())
                continue
            _cabi.call("caphn_adam_step", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                       p.numel(), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                       float(group["weight_decay"]), st["step"], gscale, _stream())
        return loss
