"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the Caption-HN hot path.

This file is the parity oracle for the CUDA path: a straight-line, functional torch-on-CPU restatement of
what the reference computes with nn.Module objects.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it -- as the checker (or the timed CPU
baseline), never as something the product path calls.  The arithmetic lives in PyTorch's CPU kernels (the reference
pins no torch version; de-facto runtime = this image's torch 2.11 CPU build), so this port calls the same torch ops
in the same order; it is differentiable, hence ``torch.autograd`` on it is also the gradient oracle.

Pinning status: the reference ships no tests/golden vectors.  This port is pinned against the
*unmodified reference modules executed in the build container* -- ``oracle/make_golden.py`` runs them through
``oracle/ref_harness.py`` on seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this port against those vectors.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, hypernet_attention.py:64,66
Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------------------------------
# head sizing rules
# ----------------------------------------------------------------------------------------------------------------------
def gru_param_shapes_attention(E: int, Fo: int, H: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """nn.GRUCell(E+F, H) parameters in named_parameters() order (models/decoderlstm.py:32)."""
    return [("weight_ih", (3 * H, E + Fo)), ("weight_hh", (3 * H, H)), ("bias_ih", (3 * H,)), ("bias_hh", (3 * H,))]


def head_dims_attention(he: int, w_size: int, N: int = 1, M: int = 500) -> Tuple[int, int, int]:
    """(in, mid, out_in) of one head, hypernet_attention.py:77-96.  Returns (he, mid, second-layer in_features)."""
    if w_size < N * he:  # :78-83 (broken branch: Linear(he,N) -> Linear(w,w); only consistent when w == N)
        return (N * he, N, w_size)
    if w_size // M < N * he:  # :85-90
        return (N * he, N * he, N * he)
    return (N * he, w_size // M, w_size // M)  # :91-96


def gru_param_shapes_pooled(E: int, H: int, L: int, gates: int = 3) -> List[Tuple[str, Tuple[int, ...]]]:
    """DecoderGRU (gates = 3) / DecoderRNN (LSTM, gates = 4) named_parameters() minus embed/fc_out, in order
    (later.py:376-381 / 241-244, hypernet.py:62-68)."""
    G = gates
    out = [("lstm_cell.weight_ih", (G * H, E)), ("lstm_cell.weight_hh", (G * H, H)),
           ("lstm_cell.bias_ih", (G * H,)), ("lstm_cell.bias_hh", (G * H,))]
    for l in range(L - 1):
        out += [(f"layers.{l}.weight_ih", (G * H, H)), (f"layers.{l}.weight_hh", (G * H, H)),
                (f"layers.{l}.bias_ih", (G * H,)), (f"layers.{l}.bias_hh", (G * H,))]
    return out


def head_dims_pooled(E: int, w_size: int) -> Tuple[int, int, int]:
    """hypernet.py:70-89."""
    if w_size < 8 * E:
        return (8 * E, w_size, w_size)
    if w_size // 8 < 8 * E:
        return (8 * E, 8 * E, 8 * E)
    return (8 * E, w_size // 8, w_size // 8)


# ----------------------------------------------------------------------------------------------------------------------
# hypernetwork: style/domain vector -> flat theta
# ----------------------------------------------------------------------------------------------------------------------
def hypernet_theta(p: Params, x: torch.Tensor, n_heads: int, prefix: str = "") -> torch.Tensor:
    """theta = cat_i flatten(head_i(hn_base(x)))   -- hypernet_attention.py:111-118, hypernet.py:104-111.

    ``x`` is ``[he]``, ``[1, he]`` or (grouped generalisation) ``[G, he]``; returns ``[theta]`` for the first two
    (exactly the reference: flatten + cat on dim 0) and ``[G, theta]`` for G > 1.
    """
    g = lambda k: p[prefix + k]
    b = F.leaky_relu(F.linear(x, g("hn_base.0.weight"), g("hn_base.0.bias")), LEAKY_SLOPE)
    b = F.leaky_relu(F.linear(b, g("hn_base.2.weight"), g("hn_base.2.bias")), LEAKY_SLOPE)
    outs = []
    for i in range(n_heads):
        a = F.leaky_relu(F.linear(b, g(f"hn_heads.{i}.0.weight"), g(f"hn_heads.{i}.0.bias")), LEAKY_SLOPE)
        outs.append(F.linear(a, g(f"hn_heads.{i}.2.weight"), g(f"hn_heads.{i}.2.bias")))
    if x.dim() == 2 and x.shape[0] > 1:
        return torch.cat(outs, dim=1)
    return torch.cat([o.flatten() for o in outs], dim=0)


def split_theta_attention(theta: torch.Tensor, E: int, Fo: int, H: int):
    """theta -> (W_ih[3H,E+F], W_hh[3H,H], b_ih, b_hh); slicing of utils.py:44-60 for a bare GRUCell."""
    out, a = [], 0
    for _, shp in gru_param_shapes_attention(E, Fo, H):
        n = int(np.prod(shp))
        out.append(theta[a:a + n].reshape(shp))
        a += n
    assert a == theta.numel()
    return tuple(out)


def split_theta_pooled(theta: torch.Tensor, E: int, H: int, L: int, gates: int = 3):
    """Per-cell weights of DecoderGRU from theta -- utils.py:44-69.

    ``count`` restarts at 0 inside every recursive call (utils.py:45), and the return value of the child call is the
    child's own count; the child ``layers`` ModuleList holds no parameters itself and recurses into ``layers.0`` ...,
    each of which again starts at offset 0.  So every extra GRUCell(H,H) reads theta from offset 0.
    """
    cells = []
    a = 0
    cell = []
    for _, shp in gru_param_shapes_pooled(E, H, 1, gates):
        n = int(np.prod(shp))
        cell.append(theta[a:a + n].reshape(shp))
        a += n
    cells.append(tuple(cell))
    for _ in range(L - 1):
        a, cell = 0, []
        for shp in ((gates * H, H), (gates * H, H), (gates * H,), (gates * H,)):
            n = int(np.prod(shp))
            cell.append(theta[a:a + n].reshape(shp))
            a += n
        cells.append(tuple(cell))
    return cells


# ----------------------------------------------------------------------------------------------------------------------
# cells
# ----------------------------------------------------------------------------------------------------------------------
def gru_cell(x, h, W_ih, W_hh, b_ih, b_hh):
    """torch nn.GRUCell, gate order r,z,n (called at models/decoderlstm.py:100, later.py:411)."""
    gi = F.linear(x, W_ih, b_ih)
    gh = F.linear(h, W_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def bahdanau(p: Params, f: torch.Tensor, h: torch.Tensor, pre: str = "captioner.attention."):
    """models/attention.py:21-46 (W_a(features) recomputed per call exactly as the reference does)."""
    a1 = F.linear(f, p[pre + "W_a.weight"], p[pre + "W_a.bias"])
    a2 = F.linear(h.unsqueeze(1), p[pre + "U_a.weight"], p[pre + "U_a.bias"])
    s = F.linear(torch.tanh(a1 + a2), p[pre + "v_a.weight"], p[pre + "v_a.bias"])
    alpha = F.softmax(s, dim=1)
    ctx = torch.sum(alpha * f, dim=1)
    return ctx, alpha.squeeze(2)


# ----------------------------------------------------------------------------------------------------------------------
# Variant B decoder: AttentionGru.forward  (models/decoderlstm.py:49-120)
# ----------------------------------------------------------------------------------------------------------------------
def attention_gru_forward(p: Params, gru_w, features: torch.Tensor, captions: torch.Tensor, sample_prob: float = 0.0,
                          rng: Optional[np.random.RandomState] = None, pre: str = "captioner."):
    """Returns (logits[B,T,V], attn[B,T,P]).

    ``gru_w`` = (W_ih, W_hh, b_ih, b_hh) generated by the hypernet.  ``rng`` stands in for NumPy's global RNG:
    one ``random()`` draw per step, also at t = 0 where it is unused (models/decoderlstm.py:79-80).
    Teacher-forced inputs are zero at t = 0 *and* t = 1 (in-place aliasing, :83-88); sampled steps feed back
    ``argmax log_softmax(logits/0.5)`` (:91-96).
    """
    rng = rng if rng is not None else np.random
    f = F.linear(features, p[pre + "feature_fc.0.weight"], p[pre + "feature_fc.0.bias"])
    f = F.linear(F.relu(f), p[pre + "feature_fc.2.weight"], p[pre + "feature_fc.2.bias"])  # :61
    emb_w = p[pre + "embed.weight"]
    embed = F.embedding(captions, emb_w).clone()  # :62
    h = F.linear(f.mean(dim=1), p[pre + "init_h.weight"], p[pre + "init_h.bias"])  # :63,133-134
    B, T = captions.shape
    outs, atts = [], []
    embed_zeroed = embed.clone()
    embed_zeroed[:, 0, :] = 0  # :83-84 zeroes embed[:,0,:] in place; t = 1 then reads that slice (:88)
    output = None
    for t in range(T):
        sp = 0.0 if t == 0 else sample_prob
        use_sampling = rng.random() < sp
        if not use_sampling:
            word = embed_zeroed[:, 0, :] if t == 0 else embed_zeroed[:, t - 1, :]
        else:
            scoring = F.log_softmax(output / 0.5, dim=1)
            top = scoring.topk(1)[1]
            word = F.embedding(top, emb_w).squeeze(1)
        ctx, alpha = bahdanau(p, f, h, pre + "attention.")
        h = gru_cell(torch.cat([word, ctx], 1), h, *gru_w)
        output = F.linear(h, p[pre + "fc.weight"], p[pre + "fc.bias"])  # dropout p=0 -> identity (:104-105)
        outs.append(output)
        atts.append(alpha)
    return torch.stack(outs, 1), torch.stack(atts, 1)


def attention_gru_greedy_search(p: Params, gru_w, features_fc: torch.Tensor, end_sentence: int = 2,
                                max_sentence: int = 20, pre: str = "captioner."):
    """AttentionGru.greedy_search (models/decoderlstm.py:138-175), B = 1: ``features_fc`` has ALREADY been through
    feature_fc; the first input is Emb[0] (not zeros); stops after ``max_sentence`` tokens or right after ``end_sentence``.
    Returns (tokens list[int], attention weights [len, P])."""
    f = features_fc
    emb_w = p[pre + "embed.weight"]
    h = F.linear(f.mean(dim=1), p[pre + "init_h.weight"], p[pre + "init_h.bias"])
    word = torch.tensor([0])
    sentence, weights = [], []
    while True:
        ctx, alpha = bahdanau(p, f, h, pre + "attention.")
        h = gru_cell(torch.cat([F.embedding(word, emb_w), ctx], 1), h, *gru_w)
        out = F.linear(h, p[pre + "fc.weight"], p[pre + "fc.bias"])
        top = F.log_softmax(out, dim=1)[0].topk(1)[1]
        sentence.append(int(top.item()))
        weights.append(alpha.reshape(-1))
        word = top
        if len(sentence) >= max_sentence or int(top.item()) == end_sentence:
            break
    return sentence, torch.stack(weights, 0)


def attention_beam_search(p: Params, gru_w, features: torch.Tensor, beam_size: int = 3, end_sentence: int = 2,
                          max_steps: int = 50, pre: str = "captioner."):
    """Beam search of HyperNet.test_step (hypernet_attention.py:247-326) for ONE image: ``features`` [1,P,D] are the
    encoder features, feature_fc is applied here (:249).  Returns the best complete sequence as a list of token ids
    (leading 0, trailing </s>), or None when the reference computes no beam caption (some beam still open after
    ``max_steps`` + 1 steps: ``compute = False`` at :311).

    Kept quirks: all k rows start from word 0 and the embeddings of EVERY row are zeroed whenever the first row's
    previous word is 0 (:265-266); step 1 ranks scores[0] only (:274-275); scores are summed log-probabilities with no
    length normalisation; finished beams are removed and k shrinks (:296-303).
    """
    emb_w = p[pre + "embed.weight"]
    V = p[pre + "fc.weight"].shape[0]
    enc = F.linear(F.relu(F.linear(features, p[pre + "feature_fc.0.weight"], p[pre + "feature_fc.0.bias"])),
                   p[pre + "feature_fc.2.weight"], p[pre + "feature_fc.2.bias"])
    k = beam_size
    enc = enc.expand(k, enc.shape[1], enc.shape[2])
    prev = torch.zeros(k, dtype=torch.long)
    seqs = prev.unsqueeze(1)
    top_scores = torch.zeros(k, 1)
    complete, complete_scores = [], []
    h = F.linear(enc.mean(dim=1), p[pre + "init_h.weight"], p[pre + "init_h.bias"])
    step = 1
    while True:
        emb = F.embedding(prev, emb_w)
        if int(prev[0]) == 0:
            emb = torch.zeros_like(emb)
        ctx, _ = bahdanau(p, enc, h, pre + "attention.")
        h = gru_cell(torch.cat([emb, ctx], 1), h, *gru_w)
        scores = F.log_softmax(F.linear(h, p[pre + "fc.weight"], p[pre + "fc.bias"]), dim=1)
        scores = top_scores.expand_as(scores) + scores
        if step == 1:
            top_scores, top_words = scores[0].topk(k, 0, True, True)
        else:
            top_scores, top_words = scores.reshape(-1).topk(k, 0, True, True)
        prev_inds = torch.div(top_words, V, rounding_mode="floor")
        next_inds = top_words % V
        seqs = torch.cat([seqs[prev_inds], next_inds.unsqueeze(1)], dim=1)
        incomplete = [i for i, w in enumerate(next_inds.tolist()) if w != end_sentence]
        done = [i for i in range(len(next_inds)) if i not in incomplete]
        if done:
            complete.extend(seqs[done].tolist())
            complete_scores.extend(top_scores[done].tolist())
        k -= len(done)
        if k == 0:
            break
        seqs = seqs[incomplete]
        h = h[prev_inds[incomplete]]
        enc = enc[prev_inds[incomplete]]
        top_scores = top_scores[incomplete].unsqueeze(1)
        prev = next_inds[incomplete]
        if step > max_steps:
            return None
        step += 1
    return complete[complete_scores.index(max(complete_scores))]


def lstm_cell(x, h, c, W_ih, W_hh, b_ih, b_hh):
    """torch nn.LSTMCell, gate order i,f,g,o (called at later.py:277-291, 344-351)."""
    g = F.linear(x, W_ih, b_ih) + F.linear(h, W_hh, b_hh)
    i, f, gg, o = g.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def caption_loss(logits: torch.Tensor, captions: torch.Tensor, ignore_index: Optional[int] = 0) -> torch.Tensor:
    """cc_train_hypernet.py:153 (ignore_index=<pad>=0) / hypernet.py:145 (ignore_index=None)."""
    V = logits.shape[-1]
    if ignore_index is None:
        return F.cross_entropy(logits.reshape(-1, V), captions.reshape(-1))
    return F.cross_entropy(logits.reshape(-1, V), captions.reshape(-1), ignore_index=ignore_index)


# ----------------------------------------------------------------------------------------------------------------------
# Variant A decoder: DecoderGRU.forward / infer  (later.py:389-490)
# ----------------------------------------------------------------------------------------------------------------------
def decoder_gru_forward(p: Params, cells, features: torch.Tensor, captions: torch.Tensor, h0: torch.Tensor,
                        pre: str = "captioner."):
    """Teacher-forced DecoderGRU.forward (later.py:389-457); ``h0`` replaces torch.rand(B,H) (:393).

    t = 0 feeds ``features`` (:411), t >= 1 feeds Emb[caps[:,t-1]] (:418); every extra layer is applied as
    ``h = layer(h, h)`` (:413-414, :420-421).  Returns logits [B,T,V].
    """
    emb = F.embedding(captions, p[pre + "embed.weight"])
    h = h0
    outs = []
    for t in range(captions.shape[1]):
        x = features if t == 0 else emb[:, t - 1, :]
        h = gru_cell(x, h, *cells[0])
        for c in cells[1:]:
            h = gru_cell(h, h, *c)
        outs.append(F.linear(h, p[pre + "fc_out.weight"], p[pre + "fc_out.bias"]))
    return torch.stack(outs, 1)


def decoder_gru_infer(p: Params, cells, features: torch.Tensor, max_len: int, h0: torch.Tensor,
                      pre: str = "captioner."):
    """DecoderGRU.infer (later.py:459-490): greedy argmax feedback, first cell only, returns softmax probs."""
    h = h0
    outs = []
    out = None
    for t in range(max_len):
        if t == 0:
            h = gru_cell(features, h, *cells[0])
        else:
            words = torch.argmax(out, dim=1)
            h = gru_cell(F.embedding(words, p[pre + "embed.weight"]), h, *cells[0])
        out = F.softmax(F.linear(h, p[pre + "fc_out.weight"], p[pre + "fc_out.bias"]), dim=1)
        outs.append(out)
    return torch.stack(outs, 1)


def decoder_rnn_forward(p: Params, cells, features: torch.Tensor, captions: torch.Tensor, pre: str = "captioner."):
    """Teacher-forced DecoderRNN.forward (later.py:254-324): (h, c) start at zero (:256-259); t = 0 feeds ``features``
    (:277), t >= 1 feeds Emb[caps[:,t-1]] (:284); every extra cell is applied as ``(h, c) = layer(h, (h, c))``
    (:279-281, :286-288).  Returns logits [B,T,V]."""
    emb = F.embedding(captions, p[pre + "embed.weight"])
    B, H = features.shape[0], cells[0][1].shape[1]
    h, c = torch.zeros(B, H), torch.zeros(B, H)
    outs = []
    for t in range(captions.shape[1]):
        x = features if t == 0 else emb[:, t - 1, :]
        h, c = lstm_cell(x, h, c, *cells[0])
        for cl in cells[1:]:
            h, c = lstm_cell(h, h, c, *cl)
        outs.append(F.linear(h, p[pre + "fc_out.weight"], p[pre + "fc_out.bias"]))
    return torch.stack(outs, 1)


def decoder_rnn_infer(p: Params, cells, features: torch.Tensor, max_len: int, pre: str = "captioner."):
    """DecoderRNN.infer (later.py:326-360): greedy argmax feedback, first cell only, returns softmax probs."""
    B, H = features.shape[0], cells[0][1].shape[1]
    h, c = torch.zeros(B, H), torch.zeros(B, H)
    outs, out = [], None
    for t in range(max_len):
        x = features if t == 0 else F.embedding(torch.argmax(out, dim=1), p[pre + "embed.weight"])
        h, c = lstm_cell(x, h, c, *cells[0])
        out = F.softmax(F.linear(h, p[pre + "fc_out.weight"], p[pre + "fc_out.bias"]), dim=1)
        outs.append(out)
    return torch.stack(outs, 1)


# ----------------------------------------------------------------------------------------------------------------------
# whole-path helpers (hypernet + decoder), used by tests and by bench.py's CPU baseline
# ----------------------------------------------------------------------------------------------------------------------
def dims_attention(p: Params, pre: str = ""):
    E = p[pre + "captioner.embed.weight"].shape[1]
    Fo = p[pre + "captioner.feature_fc.2.weight"].shape[0]
    H = p[pre + "captioner.init_h.weight"].shape[0]
    return E, Fo, H


def path_attention(p: Params, style: torch.Tensor, features, captions, sample_prob=0.0, rng=None, flow=True):
    """hypernet_attention.HyperNet.forward + AttentionGru.forward.  ``flow=False`` reproduces the reference's graph
    cut (utils.py:57 wraps every slice in nn.Parameter, detaching it); the numbers are identical either way."""
    E, Fo, H = dims_attention(p)
    theta = hypernet_theta(p, style, 4)
    gw = split_theta_attention(theta, E, Fo, H)
    if not flow:
        gw = tuple(w.detach().requires_grad_(True) for w in gw)
    logits, att = attention_gru_forward(p, gw, features, captions, sample_prob, rng)
    return logits, att, theta, gw


def path_pooled(p: Params, style: torch.Tensor, pooled, captions, h0, L=1, flow=True, infer_len=None, cell="gru"):
    """hypernet.HyperNet.forward + image_encoder.fc (hypernet.py:46,134) + DecoderGRU.forward / infer, or -- with
    ``cell="lstm"`` (hypernet.py:53, type != 'gru') -- DecoderRNN.forward / infer (``h0`` unused: zero initial state)."""
    E = p["captioner.embed.weight"].shape[1]
    H = p["captioner.fc_out.weight"].shape[1]
    gates = 3 if cell == "gru" else 4
    theta = hypernet_theta(p, style, 4 * L)
    cells = split_theta_pooled(theta, E, H, L, gates)
    if not flow:
        cells = [tuple(w.detach().requires_grad_(True) for w in c) for c in cells]
    feats = F.linear(pooled, p["image_encoder.fc.weight"], p["image_encoder.fc.bias"])
    if cell != "gru":
        if infer_len is not None:
            return decoder_rnn_infer(p, cells, feats, infer_len), theta, cells
        return decoder_rnn_forward(p, cells, feats, captions), theta, cells
    if infer_len is not None:
        return decoder_gru_infer(p, cells, feats, infer_len, h0), theta, cells
    return decoder_gru_forward(p, cells, feats, captions, h0), theta, cells


# ----------------------------------------------------------------------------------------------------------------------
# seeded synthetic inputs and default-init parameters; shared by tests and bench
# ----------------------------------------------------------------------------------------------------------------------
def synth_captions(B: int, T: int, V: int, gen: torch.Generator) -> torch.Tensor:
    """caps[b,0]=<s>=1, body ~ U{7..V-1}, caps[b,L-1]=</s>=2, 0-padded; L ~ clip(round(N(12.5,4)),4,T); row 0 full."""
    L = torch.clamp(torch.round(torch.normal(12.5, 4.0, (B,), generator=gen)), 4, T).long()
    L[0] = T
    caps = torch.randint(7, V, (B, T), generator=gen)
    caps[:, 0] = 1
    idx = torch.arange(T).unsqueeze(0)
    caps[idx == (L - 1).unsqueeze(1)] = 2
    caps[idx >= L.unsqueeze(1)] = 0
    return caps


def _linear_init(out_f, in_f, gen):
    """nn.Linear default init: U(-1/sqrt(in), 1/sqrt(in)) for weight (kaiming_uniform a=sqrt(5)) and bias."""
    k = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * k
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * k
    return w, b


def init_params_attention(D, Fo, E, H, V, he, seed=0) -> Params:
    """Random parameters with the reference state_dict layout and torch-default init distributions."""
    gen = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, o, i):
        p[name + ".weight"], p[name + ".bias"] = _linear_init(o, i, gen)

    lin("captioner.feature_fc.0", Fo, D)
    lin("captioner.feature_fc.2", Fo, Fo)
    p["captioner.embed.weight"] = torch.randn(V, E, generator=gen)
    lin("captioner.fc", V, H)
    lin("captioner.attention.W_a", H, Fo)
    lin("captioner.attention.U_a", H, H)
    lin("captioner.attention.v_a", 1, H)
    lin("captioner.init_h", H, Fo)
    lin("hn_base.0", he, he)
    lin("hn_base.2", he, he)
    for i, (_, shp) in enumerate(gru_param_shapes_attention(E, Fo, H)):
        w = int(np.prod(shp))
        i0, mid, i2 = head_dims_attention(he, w)
        lin(f"hn_heads.{i}.0", mid, i0)
        lin(f"hn_heads.{i}.2", w, i2)
    return p


def init_params_pooled(D, E, H, V, L=1, seed=0, gates=3) -> Params:
    gen = torch.Generator().manual_seed(seed)
    p: Params = {}

    def lin(name, o, i):
        p[name + ".weight"], p[name + ".bias"] = _linear_init(o, i, gen)

    lin("image_encoder.fc", E, D)
    lin("captioner.fc_out", V, H)
    p["captioner.embed.weight"] = torch.randn(V, E, generator=gen)
    lin("hn_base.0", 4 * E, E)
    lin("hn_base.2", 8 * E, 4 * E)
    for i, (_, shp) in enumerate(gru_param_shapes_pooled(E, H, L, gates)):
        w = int(np.prod(shp))
        i0, mid, i2 = head_dims_pooled(E, w)
        lin(f"hn_heads.{i}.0", mid, i0)
        lin(f"hn_heads.{i}.2", w, i2)
    return p
