"""Per-style grouped decoding (G style groups in one batch): one hypernet weight pass for all groups; oracle = one
reference-semantics call per group, concatenated, loss over the concatenation (SURVEY.md 8(c) "grouped oracle")."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import rel_err, grad_close
from oracle import caption_hn_oracle as O


def _load(m, p):
    sd = m.state_dict(); sd.update(p); m.load_state_dict(sd)
    return m.cuda()


@pytest.mark.parametrize("G", [3, 11])
def test_attention_grouped_matches_per_group_oracle(G):
    import hypernet_image_captioning_b200 as C
    B, T, Fo, E, H, V = 14, 6, 16, 12, 20, 70
    p = O.init_params_attention(2048, Fo, E, H, V, E, seed=2)
    g = torch.Generator().manual_seed(8)
    feats = torch.randn(B, 49, 2048, generator=g)
    caps = O.synth_captions(B, T, V, g)
    styles = torch.randn(G, E, generator=g)
    groups = torch.randint(0, G, (B,), generator=g)
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    logits_ref = torch.zeros(B, T, V)
    rows = []
    for gi in range(G):
        idx = (groups == gi).nonzero().squeeze(1)
        if idx.numel() == 0:
            continue
        lg, _, _, _ = O.path_attention(pl, styles[gi:gi + 1], feats[idx], caps[idx], 0.0, np.random.RandomState(0))
        rows.append((idx, lg))
    logits_ref = torch.cat([lg for _, lg in rows], 0)[torch.argsort(torch.cat([i for i, _ in rows]))]
    O.caption_loss(logits_ref, caps, 0).backward()

    m = _load(C.HyperNetAttention(Fo, E, H, V, None), p)
    captioner = m.forward_grouped(styles.cuda())
    logits, att = captioner(feats.cuda(), caps.cuda(), 0.0, groups=groups.cuda())
    C.cross_entropy(logits, caps.cuda(), 0).backward()
    assert rel_err(logits, logits_ref) < 1e-4
    for k, v in m.named_parameters():
        if k.startswith("captioner.gru."):
            continue
        assert grad_close(v.grad, pl[k].grad, 1e-3), k


def test_pooled_grouped_matches_per_group_oracle():
    import hypernet_image_captioning_b200 as C
    G, B, T, E, H, V = 3, 10, 5, 16, 12, 97
    p = O.init_params_pooled(2048, E, H, V, seed=4)
    g = torch.Generator().manual_seed(9)
    pooled = torch.relu(torch.randn(B, 2048, generator=g))
    caps = O.synth_captions(B, T, V, g)
    styles = torch.randn(G, E, generator=g)
    h0 = torch.rand(B, H, generator=g)
    groups = torch.tensor([0, 1, 2, 0, 1, 2, 2, 2, 0, 1])
    pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    outs, idxs = [], []
    for gi in range(G):
        idx = (groups == gi).nonzero().squeeze(1)
        lg, _, _ = O.path_pooled(pl, styles[gi:gi + 1], pooled[idx], caps[idx], h0[idx])
        outs.append(lg); idxs.append(idx)
    logits_ref = torch.cat(outs, 0)[torch.argsort(torch.cat(idxs))]
    O.caption_loss(logits_ref, caps, None).backward()

    m = _load(C.HyperNetPooled(E, H, V, None), p)
    captioner = m.forward_grouped(styles.cuda())
    logits = captioner(m.image_encoder(pooled.cuda()), caps.cuda(), True, h0=h0.cuda(), groups=groups.cuda())
    C.cross_entropy(logits, caps.cuda(), None).backward()
    assert rel_err(logits, logits_ref) < 1e-4
    for k, v in m.named_parameters():
        if k.startswith("captioner.lstm_cell."):
            continue
        assert grad_close(v.grad, pl[k].grad, 1e-3), k
