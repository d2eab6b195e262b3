"""Diagnostic (GPU): per-parameter gradient error of the full-size attention step vs the oracle, with the side streams on
and off, plus the weight-gradient GEMM of feature_fc.0 (M=200, N=2048, K=B*P) in isolation against fp64."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hypernet_image_captioning_b200 as C  # noqa: E402
from hypernet_image_captioning_b200 import ops, streams  # noqa: E402
from oracle import caption_hn_oracle as O  # noqa: E402
from golden_util import rel_err  # noqa: E402


def gemm_check(K, M=200, N=2048, sparse=True, scale=1.0):
    g = torch.Generator().manual_seed(K)
    A = torch.randn(K, M, generator=g) * scale
    if sparse:
        A = A * (torch.rand(K, M, generator=g) > 0.5)
    Bm = torch.randn(K, N, generator=g)
    ref = (A.double().t() @ Bm.double())
    Ad, Bd = A.cuda(), Bm.cuda()
    out = ops.gemm_tc(ops.split_bf16(Ad, mn=True), ops.split_bf16(Bd, mn=True))
    out2 = ops.gemm_tc(ops.split_bf16_t(Ad), ops.split_bf16_t(Bd))
    t32 = (Ad.t() @ Bd)
    torch.cuda.synchronize()
    e = (out.double().cpu() - ref).abs()
    print(f"gemm K={K} M={M} N={N}: mn-major rel {rel_err(out, ref):.2e}  k-major(transposing split) {rel_err(out2, ref):.2e}  "
          f"torch fp32 {rel_err(t32, ref):.2e}; worst row {int(e.amax(1).argmax())} worst col {int(e.amax(0).argmax())} "
          f"row-block err {[f'{x:.1e}' for x in (e.amax(1).reshape(-1, 50).amax(1) / ref.abs().max()).tolist()]}")


def step_check(overlap, async_hn, B=512, fused=False):
    d = dict(B=B, T=20, Fo=200, E=200, H=200, V=9684, P=49, D=2048)
    p = O.init_params_attention(d["D"], d["Fo"], d["E"], d["H"], d["V"], d["E"], seed=11)
    g = torch.Generator().manual_seed(21)
    feats = torch.randn(B, d["P"], d["D"], generator=g)
    caps = O.synth_captions(B, d["T"], d["V"], g)
    style = p["captioner.embed.weight"][4:5].clone()
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    logits_ref, att_ref, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0), flow=True)
    O.caption_loss(logits_ref, caps, 0).backward()
    streams.ENABLED = overlap
    m = C.HyperNetAttention(200, 200, 200, d["V"], None)
    sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
    m = m.cuda()
    m.async_hypernet = async_hn
    cap = m.forward(style.cuda())
    np.random.seed(0)
    if fused:
        loss, logits, att = cap.forward_loss(feats.cuda(), caps.cuda(), 0.0, ignore_index=0)
    else:
        logits, att = cap(feats.cuda(), caps.cuda(), 0.0)
        loss = C.cross_entropy(logits, caps.cuda(), 0)
    loss.backward()
    torch.cuda.synchronize()
    print(f"--- step B={B} overlap={overlap} async_hn={async_hn} fused={fused}: logits {rel_err(logits, logits_ref):.2e}")
    for k, v in m.named_parameters():
        if k.startswith("captioner.gru."):
            continue
        print(f"    {k:40s} {rel_err(v.grad, pl[k].grad):.2e}   max|ref| {pl[k].grad.abs().max().item():.2e}")


if __name__ == "__main__":
    for K in (1568, 6272, 25088):
        gemm_check(K)
    gemm_check(25088, sparse=False)
    gemm_check(25088, scale=1e-5)
    step_check(False, False, B=128)
    step_check(False, False)
    step_check(True, False)
    step_check(True, True)
    step_check(True, True, fused=True)
