"""Loading of the committed golden vectors (tests/golden/*.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = v if v.dtype.kind in "US" else torch.from_numpy(v)
    return out


def params_of(case):
    return {k[3:]: v.clone() for k, v in case.items() if k.startswith("sd/")}


def rel_err(a, b):
    """max|a-b| / max|b|  (the tolerance form of SURVEY.md 8(d))."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = (a - b).abs().max().item()
    s = b.abs().max().item()
    return d / s if s > 0 else d


def grad_close(a, b, rtol, atol=1e-7):
    """Gradient comparison: relative to max|ref|, with an absolute floor for gradients that are analytically zero
    (e.g. d/d v_a.bias -- softmax is shift invariant -- which come out as 1e-9 rounding noise)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    d = (a - b).abs().max().item()
    return d <= atol or d <= rtol * b.abs().max().item()
