"""Module-level parity of the attention path (HyperNetAttention + AttentionGru) through the drop-in API: golden vectors
of the unmodified reference (tests/golden/attn_*.npz) and the oracle port on larger seeded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import load_case, params_of, rel_err, grad_close
from oracle import caption_hn_oracle as O

TOL_LOGITS = 1e-4   # BASELINE.json north_star: fp32 mode within 1e-4 relative on logits and loss
TOL_GRAD = 1e-3


def _model_from(p, Fo, E, H, V, cc, he, mode="flow"):
    import hypernet_image_captioning_b200 as C
    m = C.HyperNetAttention(Fo, E, H, V, None, cc=cc, hyper_emb=he)
    sd = m.state_dict()
    missing = [k for k in p if k not in sd]
    assert not missing, missing
    sd.update(p)
    m.load_state_dict(sd)
    m.grad_mode = mode
    return m.cuda()


CASES = [("attn_flickr", False, 10), ("attn_cc", True, 10)]


@pytest.mark.parametrize("name,cc,he", CASES)
def test_state_dict_layout_matches_reference(name, cc, he):
    import hypernet_image_captioning_b200 as C
    c = load_case(name)
    p = params_of(c)
    m = C.HyperNetAttention(16, 12, 20, 50, None, cc=cc, hyper_emb=he)
    sd = m.state_dict()
    gen = {"captioner.gru." + k for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")}
    assert set(sd.keys()) == set(p.keys()) | gen
    for k, v in p.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


@pytest.mark.parametrize("name,cc,he", CASES)
@pytest.mark.parametrize("mode", ["literal", "flow"])
def test_attention_golden_teacher_forced(name, cc, he, mode):
    import hypernet_image_captioning_b200 as C
    c = load_case(name)
    m = _model_from(params_of(c), 16, 12, 20, 50, cc, he, mode)
    captioner = m.forward(c["style"].cuda())
    for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
        assert rel_err(getattr(captioner.gru, k), c["gen/" + k]) < 1e-5, k
    np.random.seed(0)
    logits, att = captioner(c["features"].cuda(), c["captions"].cuda(), 0.0)
    assert rel_err(logits, c["tf/logits"]) < TOL_LOGITS
    assert rel_err(att, c["tf/attn"]) < TOL_LOGITS
    loss = C.cross_entropy(logits, c["captions"].cuda(), 0)
    assert abs(loss.item() - c["tf/loss"].item()) < TOL_LOGITS * abs(c["tf/loss"].item())
    loss.backward()
    named = dict(m.named_parameters())
    for k, v in c.items():
        if not k.startswith("tf/grad/") or k.startswith("tf/grad/captioner.gru."):
            continue
        assert grad_close(named[k[8:]].grad, v, TOL_GRAD), k
    if mode == "literal":
        for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            assert grad_close(getattr(captioner.gru, k).grad, c["tf/grad/captioner.gru." + k], TOL_GRAD), k
        assert all(v.grad is None for k, v in named.items() if k.startswith("hn_"))
    else:
        n = 0
        for k, v in c.items():
            if k.startswith("flow/grad/"):
                assert grad_close(named[k[10:]].grad, v, TOL_GRAD), k
                n += 1
        assert n >= 12


@pytest.mark.parametrize("table", [False, True])
@pytest.mark.parametrize("name,cc,he", CASES)
def test_attention_golden_greedy_token_exact(name, cc, he, table, monkeypatch):
    """table = True forces the decode-time projection table (P = [Emb; 0] W_ih[:, :E]^T + b_ih, one GEMM per call) that
    large batches use instead of a per-step gather + GEMM; both must reproduce the reference's greedy tokens."""
    from hypernet_image_captioning_b200 import graphs, ops
    monkeypatch.setattr(ops, "use_projection_table", lambda B, steps, V: table)
    graphs.clear()
    c = load_case(name)
    m = _model_from(params_of(c), 16, 12, 20, 50, cc, he)
    np.random.seed(0)
    with torch.no_grad():
        captioner = m.forward(c["style"].cuda())
        logits, att = captioner(c["features"].cuda(), c["captions"].cuda(), 1.0)
    assert torch.equal(logits.argmax(-1).cpu(), c["greedy/logits"].argmax(-1))
    assert rel_err(logits, c["greedy/logits"]) < TOL_LOGITS
    assert rel_err(att, c["greedy/attn"]) < TOL_LOGITS


def test_numpy_rng_consumption_matches_reference():
    """One np.random.random() per step, also at t = 0 (models/decoderlstm.py:79-80)."""
    c = load_case("attn_flickr")
    m = _model_from(params_of(c), 16, 12, 20, 50, False, 10)
    T = c["captions"].shape[1]
    np.random.seed(123)
    with torch.no_grad():
        m.forward(c["style"].cuda())(c["features"].cuda(), c["captions"].cuda(), 0.0)
    after = np.random.random()
    np.random.seed(123)
    for _ in range(T):
        np.random.random()
    assert after == np.random.random()


@pytest.mark.parametrize("B,T,Fo,E,H,V,he,cc", [(13, 7, 24, 20, 28, 211, 20, False), (32, 20, 200, 200, 200, 1500, 100, True)])
def test_attention_vs_oracle_medium(B, T, Fo, E, H, V, he, cc):
    import hypernet_image_captioning_b200 as C
    p = O.init_params_attention(2048, Fo, E, H, V, he if cc else E, seed=7)
    g = torch.Generator().manual_seed(17)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    if cc:
        style = torch.zeros(he); style[5] = 1.0
    else:
        style = torch.randn(1, E, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    logits_ref, att_ref, _, _ = O.path_attention(pl, style, feats, caps, 0.0, np.random.RandomState(0), flow=True)
    loss_ref = O.caption_loss(logits_ref, caps, 0)
    loss_ref.backward()

    m = _model_from(p, Fo, E, H, V, cc, he)
    captioner = m.forward(style.cuda())
    logits, att = captioner(feats.cuda(), caps.cuda(), 0.0)
    loss = C.cross_entropy(logits, caps.cuda(), 0)
    loss.backward()
    assert rel_err(logits, logits_ref) < TOL_LOGITS
    assert rel_err(att, att_ref) < TOL_LOGITS
    assert abs(loss.item() - loss_ref.item()) < TOL_LOGITS * abs(loss_ref.item())
    for k, v in m.named_parameters():
        if k.startswith("captioner.gru."):
            continue
        assert grad_close(v.grad, pl[k].grad, TOL_GRAD), k

    # greedy decode on the same model: token-exact against the oracle
    with torch.no_grad():
        gl_ref, _, _, _ = O.path_attention(p, style, feats, caps, 1.0, np.random.RandomState(0))
        gl, _ = m.forward(style.cuda())(feats.cuda(), caps.cuda(), 1.0)
    match = (gl.argmax(-1).cpu() == gl_ref.argmax(-1)).float().mean().item()
    assert match == 1.0, f"greedy token match {match}"
    assert rel_err(gl, gl_ref) < TOL_LOGITS


def test_attention_scheduled_sampling_mixed_matches_oracle():
    """0 < sample_prob < 1: per-step host decisions from NumPy's RNG, gradient still flows through fed embeddings."""
    import hypernet_image_captioning_b200 as C
    B, T, Fo, E, H, V = 6, 8, 16, 12, 20, 60
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=9)
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    style = torch.randn(1, E, generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lr, _, _, _ = O.path_attention(pl, style, feats, caps, 0.5, np.random.RandomState(4), flow=True)
    O.caption_loss(lr, caps, 0).backward()
    m = _model_from(p, Fo, E, H, V, False, 10)
    np.random.seed(4)
    logits, _ = m.forward(style.cuda())(feats.cuda(), caps.cuda(), 0.5)
    C.cross_entropy(logits, caps.cuda(), 0).backward()
    assert rel_err(logits, lr) < TOL_LOGITS
    for k in ("captioner.embed.weight", "hn_heads.0.2.weight", "captioner.attention.U_a.weight"):
        assert grad_close(dict(m.named_parameters())[k].grad, pl[k].grad, TOL_GRAD), k


@pytest.mark.parametrize("B,T,Fo,E,H,V", [(32, 20, 200, 200, 200, 1500), (6, 5, 16, 12, 20, 60)])
def test_attention_fused_loss_matches_unfused(B, T, Fo, E, H, V):
    import hypernet_image_captioning_b200 as C
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=7)
    g = torch.Generator().manual_seed(17)
    feats = torch.randn(B, 49, 2048, generator=g).cuda()
    caps = O.synth_captions(B, T, V, g).cuda()
    style = torch.randn(1, E, generator=g).cuda()
    res = []
    for fused in (False, True):
        m = _model_from(p, Fo, E, H, V, False, 10)
        captioner = m.forward(style)
        np.random.seed(0)
        if fused:
            loss, logits, att = captioner.forward_loss(feats, caps, 0.0, ignore_index=0)
        else:
            logits, att = captioner(feats, caps, 0.0)
            loss = C.cross_entropy(logits, caps, 0)
        loss.backward()
        res.append((loss.item(), logits.detach(), {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}))
    assert abs(res[0][0] - res[1][0]) < 1e-6 * abs(res[0][0])
    assert torch.equal(res[0][1], res[1][1])
    for k, v in res[0][2].items():
        assert grad_close(res[1][2][k], v, TOL_GRAD), k


@pytest.mark.parametrize("B,T,Fo,E,H,P", [(5, 3, 16, 12, 20, 49), (40, 6, 200, 200, 200, 49), (33, 4, 64, 32, 100, 7)])
def test_attgru_cluster_forward_matches_streaming_kernel(B, T, Fo, E, H, P):
    """Weights-resident cluster forward (warp-MMA bf16x3 fragments in registers, DSMEM exchanges) == streaming kernel."""
    from hypernet_image_captioning_b200 import ops
    assert ops.attgru_cluster_ok(H, Fo, P, force=True)
    g = torch.Generator().manual_seed(B + T + H)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    Kp, f, GIw = r(B, P, H) * 0.5, r(B, P, Fo) * 0.5, r(T * B, 3 * H) * 0.5
    Ua, W_ih, W_hh = r(H, H) / H ** 0.5, r(3 * H, E + Fo) / (E + Fo) ** 0.5, r(3 * H, H) / H ** 0.5
    bu, va, bv, bhh, h0 = r(H) * 0.1, r(H) * 0.3, r(1), r(3 * H) * 0.1, r(B, H) * 0.5
    outs = []
    for cluster in (False, True):
        Hall = torch.empty(T + 1, B, H).cuda(); Hall[0] = h0
        Hbm, attn = torch.empty(B, T, H).cuda(), torch.empty(B, T, P).cuda()
        XC, saved = torch.zeros(T * B, E + Fo).cuda(), torch.empty(5, T, B, H).cuda()
        if cluster:
            ops.attgru_cluster_fwd(Kp, f, GIw, Ua, bu, va, bv, W_ih, W_hh, bhh, Hall, Hbm, attn, XC, E, saved, 0, T)
        else:
            lw = ops.AttGruWeights(W_ih, W_hh, Ua, E, P, step=False)
            ops.attgru_seq_fwd(Kp, f, GIw, lw, bu, va, bv, bhh, Hall, Hbm, attn, XC, E, saved, 0, T)
        torch.cuda.synchronize()
        outs.append((Hall, Hbm, attn, XC, saved))
    for name, x, y in zip(("Hall", "Hbm", "attn", "XC", "saved"), outs[1], outs[0]):
        assert rel_err(x, y) < 3e-5, name


@pytest.mark.parametrize("B,T,Fo,E,H,P", [(5, 3, 16, 12, 20, 49), (40, 6, 200, 200, 200, 49), (33, 4, 64, 32, 100, 7),
                                          (70, 3, 36, 8, 200, 5), (3, 2, 8, 5, 12, 3), (37, 3, 20, 6, 208, 4),
                                          (3, 2, 7, 5, 9, 3)])
def test_attgru_step_split_forward_matches_streaming_kernel(B, T, Fo, E, H, P):
    """Step-split path (U / attention / gates kernels, PDL-chained) == persistent streaming kernel, also one step at a
    time (the decode call pattern, workspace resumed between calls) and for ragged sizes (H, F not multiples of 16).
    The last case is a shape the step-split path does not cover (P*H not a multiple of 4): it must fall back."""
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(B + T + H)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    Kp, f, GIw = r(B, P, H) * 0.5, r(B, P, Fo) * 0.5, r(T * B, 3 * H) * 0.5
    Ua, W_ih, W_hh = r(H, H) / H ** 0.5, r(3 * H, E + Fo) / (E + Fo) ** 0.5, r(3 * H, H) / H ** 0.5
    bu, va, bv, bhh, h0 = r(H) * 0.1, r(H) * 0.3, r(1), r(3 * H) * 0.1, r(B, H) * 0.5
    outs = []
    for mode in ("stream", "step", "step1"):
        Hall = torch.empty(T + 1, B, H).cuda(); Hall[0] = h0
        Hbm, attn = torch.empty(B, T, H).cuda(), torch.empty(B, T, P).cuda()
        XC, saved = torch.zeros(T * B, E + Fo).cuda(), torch.empty(5, T, B, H).cuda()
        lw = ops.AttGruWeights(W_ih, W_hh, Ua, E, P, step=(mode != "stream"))
        covered = ops._attstep_bytes(H, Fo, P, 0)[0] > 0
        assert covered == (H % 4 == 0 and (P * Fo) % 4 == 0 and max(H, Fo) <= 208)
        assert (lw.pack is not None) == (mode != "stream" and covered)
        if mode == "step1":
            for t in range(T):
                ops.attgru_fwd(Kp, f, GIw, lw, bu, va, bv, bhh, Hall, Hbm, attn, XC, E, saved, t, t + 1)
        else:
            ops.attgru_fwd(Kp, f, GIw, lw, bu, va, bv, bhh, Hall, Hbm, attn, XC, E, saved, 0, T)
        torch.cuda.synchronize()
        outs.append((Hall, Hbm, attn, XC, saved))
    for name, x, y in zip(("Hall", "Hbm", "attn", "XC", "saved"), outs[1], outs[0]):
        assert rel_err(x, y) < 3e-5, name
    for name, x, y in zip(("Hall", "Hbm", "attn", "XC", "saved"), outs[2], outs[1]):
        assert torch.equal(x, y), name


@pytest.mark.parametrize("B,T,Fo,E,H,P,ext", [(5, 3, 16, 12, 20, 49, True), (70, 5, 200, 200, 200, 49, False),
                                              (33, 4, 64, 32, 100, 7, True), (130, 2, 36, 8, 208, 5, False),
                                              (3, 2, 8, 5, 12, 3, True)])
def test_attgru_step_split_backward_matches_streaming_kernel(B, T, Fo, E, H, P, ext):
    """Step-split BPTT (tensor-core gate / GEMM kernels, TMA-fed attention backward, deferred dK) == persistent kernel."""
    from hypernet_image_captioning_b200 import ops
    g = torch.Generator().manual_seed(B + T + H)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    Kp, f, GIw = r(B, P, H) * 0.5, r(B, P, Fo) * 0.5, r(T * B, 3 * H) * 0.5
    Ua, W_ih, W_hh = r(H, H) / H ** 0.5, r(3 * H, E + Fo) / (E + Fo) ** 0.5, r(3 * H, H) / H ** 0.5
    bu, va, bv, bhh, h0 = r(H) * 0.1, r(H) * 0.3, r(1), r(3 * H) * 0.1, r(B, H) * 0.5
    Hall = torch.empty(T + 1, B, H).cuda(); Hall[0] = h0
    Hbm, attn = torch.empty(B, T, H).cuda(), torch.empty(B, T, P).cuda()
    XC, saved = torch.zeros(T * B, E + Fo).cuda(), torch.empty(5, T, B, H).cuda()
    lw = ops.AttGruWeights(W_ih, W_hh, Ua, E, P, step=False)
    ops.attgru_fwd(Kp, f, GIw, lw, bu, va, bv, bhh, Hall, Hbm, attn, XC, E, saved, 0, T)
    dHbm = r(B, T, H) * 0.3
    dattn = r(B, T, P) * 0.2 if ext else None
    assert ops._attstep_bwd_bytes(H, Fo, P, B, T)[0] > 0
    ref = ops.attgru_bwd(dHbm, dattn, Kp, f, attn, saved, Hall, Ua, va, W_ih, W_hh, E, step=False)
    out = ops.attgru_bwd(dHbm, dattn, Kp, f, attn, saved, Hall, Ua, va, W_ih, W_hh, E, step=True)
    torch.cuda.synchronize()
    for name, x, y in zip(("dGI", "dGH", "dU", "dCTX", "dK", "dva", "dbv", "dh0"), out, ref):
        if name == "dbv":      # analytically zero (softmax is shift invariant): both are rounding noise
            assert abs(float(x) - float(y)) < 1e-5
        else:
            assert rel_err(x, y) < 5e-5, name


@pytest.mark.parametrize("name,cc,he", CASES)
def test_greedy_search_golden(name, cc, he):
    """B = 1 greedy_search with EOS stop (models/decoderlstm.py:138-175) against the reference's own output."""
    c = load_case(name)
    m = _model_from(params_of(c), 16, 12, 20, 50, cc, he)
    captioner = m.forward(c["style"].cuda())
    for bi in range(c["features"].shape[0]):
        with torch.no_grad():
            fproj = captioner.feature_fc(c["features"][bi:bi + 1].cuda())     # torch nn.Sequential, as the caller does
        toks, wts = captioner.greedy_search(fproj, end_sentence=2, max_sentence=7)
        assert toks == c[f"gs/{bi}/tokens"].tolist()
        assert rel_err(torch.cat(wts, 0), c[f"gs/{bi}/weights"]) < TOL_LOGITS


def test_regression_pretraining_loss_matches_reference_recipe():
    """train_init.py:70-123: loss = sum_i MSE(head_i(hn_base(style)), W_i) ; gradients reach heads and base."""
    import torch.nn.functional as F
    Fo, E, H, V = 16, 12, 20, 60
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=23)
    g = torch.Generator().manual_seed(4)
    style = torch.randn(1, E, generator=g)
    targets = [torch.randn(*shp, generator=g) * 0.1 for _, shp in O.gru_param_shapes_attention(E, Fo, H)]
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    theta = O.hypernet_theta(pl, style, 4)
    loss_ref, a = 0.0, 0
    for t in targets:
        loss_ref = loss_ref + F.mse_loss(theta[a:a + t.numel()], t.flatten())
        a += t.numel()
    loss_ref.backward()
    m = _model_from(p, Fo, E, H, V, False, 10)
    loss = m.regression_loss(style.cuda(), [t.cuda() for t in targets])
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * abs(loss_ref.item())
    for k, v in m.named_parameters():
        if k.startswith("hn_"):
            assert grad_close(v.grad, pl[k].grad, TOL_GRAD), k


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "noeos"])
def test_beam_search_golden(tag):
    """Beam search k = 3 (hypernet_attention.py:247-326) against the caption the unmodified reference's test_step produced."""
    c = load_case("attn_flickr")
    p = params_of(c)
    p["captioner.fc.bias"] = c[f"beam/{tag}/fc_bias"]
    p["captioner.fc.weight"] = c[f"beam/{tag}/fc_weight"]
    m = _model_from(p, 16, 12, 20, 50, False, 10)
    style = m.captioner.embed.weight.detach()[int(c["beam/style_id"])].reshape(1, -1)     # :244-246
    captioner = m.forward(style)
    for bi in range(c["features"].shape[0]):
        got = captioner.beam_search(c["features"][bi:bi + 1].cuda(), beam_size=3, end_sentence=2, max_steps=50)
        want = c[f"beam/{tag}/{bi}"].tolist()
        assert (got if got is not None else [-1]) == want, (tag, bi)


def test_greedy_projection_table_equals_per_step_path_at_bench_size(monkeypatch):
    """B = 512, T = 20, V = 9684: the table path (chosen automatically at this size) feeds back the same tokens and gives
    the same logits as the per-step gather + GEMM path."""
    import hypernet_image_captioning_b200 as C
    from hypernet_image_captioning_b200 import graphs, ops
    from hypernet_image_captioning_b200.synth import synth_captions
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = C.HyperNetAttention(64, 48, 56, 9684, None)
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(512, 49, 2048, generator=g).cuda()
    caps = synth_captions(512, 20, 9684, g).cuda()
    assert ops.use_projection_table(512, 21, 9684)
    outs = []
    for table in (False, True):
        monkeypatch.setattr(ops, "use_projection_table", lambda B, steps, V, table=table: table)
        graphs.clear()
        np.random.seed(0)
        with torch.no_grad():
            cap = m.forward(m.captioner.embed.weight[4:5])
            logits, att = cap(feats, caps, 1.0)
        outs.append((logits.clone(), att.clone()))
    graphs.clear()
    from golden_util import same_greedy_paths
    ok, why = same_greedy_paths(outs[0][0], outs[1][0])       # tokens compared up to the first near-tie (random-init model)
    assert ok, why
    same = (outs[0][0].argmax(-1) == outs[1][0].argmax(-1)).all(1)
    assert same.any() and rel_err(outs[1][1][same], outs[0][1][same]) < 1e-4


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "noeos"])
def test_batched_beam_search_golden(tag):
    """Device-resident batched beam search (all images of the golden case in ONE call) == the captions the unmodified
    reference's test_step produced image by image (hypernet_attention.py:247-326)."""
    c = load_case("attn_flickr")
    p = params_of(c)
    p["captioner.fc.bias"] = c[f"beam/{tag}/fc_bias"]
    p["captioner.fc.weight"] = c[f"beam/{tag}/fc_weight"]
    m = _model_from(p, 16, 12, 20, 50, False, 10)
    style = m.captioner.embed.weight.detach()[int(c["beam/style_id"])].reshape(1, -1)
    captioner = m.forward(style)
    feats = c["features"].cuda()
    got = captioner.beam_search_batched(feats, beam_size=3, end_sentence=2, max_steps=50)
    for bi in range(feats.shape[0]):
        want = c[f"beam/{tag}/{bi}"].tolist()
        assert (got[bi] if got[bi] is not None else [-1]) == want, (tag, bi)
    # a batch with repeated / reordered images gives the same per-image answers (rows of different images never mix)
    idx = torch.tensor([1, 0, 1, 2, 0][:max(2, feats.shape[0] + 2)]) % feats.shape[0]
    got2 = captioner.beam_search_batched(feats[idx], beam_size=3, end_sentence=2, max_steps=50, sync_every=1)
    for j, bi in enumerate(idx.tolist()):
        assert got2[j] == got[bi]


def test_batched_beam_search_matches_oracle_at_scale():
    """B = 48 images, V = 1500, F = E = H = 200 (step-split kernels, tensor-core GEMMs), k = 3 and 5: every image's caption
    equals the oracle's per-image beam search (which restates the reference loop).  A beam decision between two candidates
    whose scores differ by less than the 1e-5-class rounding of the bf16x3 products is numerically ambiguous (about one
    decision in a thousand on a random-init model); such images are identified by re-running the ORACLE with the vocabulary
    weights perturbed at 2e-5 relative and are excluded -- all others must match token for token."""
    B, Fo, E, H, V = 48, 200, 200, 200, 1500
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=31)
    p["captioner.fc.weight"][2] *= 60.0                               # </s> logit swings with h: beams finish at different steps
    p["captioner.fc.bias"][2] -= 1.0
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, 49, 2048, generator=g)
    style = torch.randn(1, E, generator=g)
    m = _model_from(p, Fo, E, H, V, False, 10)
    captioner = m.forward(style.cuda())
    theta = O.hypernet_theta(p, style, 4)
    gw = O.split_theta_attention(theta, E, Fo, H)
    variants = [p]
    for i in range(3):
        q = dict(p)
        q["captioner.fc.weight"] = p["captioner.fc.weight"] * (1 + 2e-5 * torch.randn(V, H, generator=g))
        variants.append(q)
    for k in (3, 5):
        got = captioner.beam_search_batched(feats.cuda(), beam_size=k, end_sentence=2, max_steps=20)
        n_checked = n_none = 0
        for bi in range(B):
            wants = [O.attention_beam_search(q, gw, feats[bi:bi + 1], beam_size=k, end_sentence=2, max_steps=20)
                     for q in variants]
            if any(w != wants[0] for w in wants[1:]):
                continue                                             # a near-tie somewhere along this image's search
            assert got[bi] == wants[0], (k, bi, got[bi], wants[0])
            n_checked += 1
            n_none += wants[0] is None
        print(f"[beam k={k}] {n_checked}/{B} images unambiguous and identical ({n_none} without a caption), "
              f"lengths {sorted(len(x) for x in got if x)[::8]}")
        assert n_checked >= 0.8 * B
