"""Token-level caption metrics on the device (csrc/metrics.cu via hypernet_image_captioning_b200.metrics) against the
reference's cap_to_text / cap_to_text_gt outputs (tests/golden/metrics.npz) and the oracle's compute_bleu restatement."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from golden_util import load_case
from oracle import metrics_oracle as MO


def _logits(c):
    B, T = c["pred_ids"].shape
    logits = torch.zeros(B, T, int(c["V"]))
    logits.scatter_(2, c["pred_ids"].unsqueeze(-1), 1.0)
    logits.scatter_(2, c["tie_ids"].unsqueeze(-1), 1.0)
    return logits


def test_caption_tokens_golden():
    from hypernet_image_captioning_b200 import metrics
    c = load_case("metrics")
    hyp, hl = metrics.caption_tokens(_logits(c).cuda())
    assert hyp.dtype == torch.int64 and hl.dtype == torch.int32
    assert torch.equal(hyp.cpu(), c["hyp"]) and torch.equal(hl.cpu(), c["hyp_len"])
    ref, rl = metrics.caption_tokens(c["gt_ids"].cuda())
    assert torch.equal(ref.cpu(), c["ref"]) and torch.equal(rl.cpu(), c["ref_len"])
    ref2, rl2 = metrics.caption_tokens(c["gt_ids"].float().cuda())     # the loaders hand captions over as float
    assert torch.equal(ref2, ref) and torch.equal(rl2, rl)


def test_tokens_to_text_one_copy():
    from types import SimpleNamespace
    from hypernet_image_captioning_b200 import metrics
    c = load_case("metrics")
    voc = SimpleNamespace(i2w={i: f"w{i}" for i in range(int(c["V"]))})
    hyp, hl = metrics.caption_tokens(_logits(c).cuda())
    text = metrics.tokens_to_text(hyp, hl, voc)
    for b, s in enumerate(text):
        assert s == " ".join(f"w{i}" for i in c["hyp"][b, :int(c["hyp_len"][b])].tolist())
    assert metrics.tokens_to_text(hyp, hl, voc, tokenized=True)[1] == []


@pytest.mark.parametrize("B,T,vocab,lo", [(8, 12, 9684, 0), (64, 20, 12, 0), (513, 40, 5, 0), (3, 1, 4, 0),
                                           (256, 64, 4, 3)])
def test_bleu_counts_vs_oracle(B, T, vocab, lo):
    """Small vocabularies give repeated n-grams (clipping) and ties; includes empty captions and T = 1; lo = 3 keeps
    the special ids out, so every caption is T = 64 tokens of a 4-word vocabulary (maximum n-gram repetition)."""
    from hypernet_image_captioning_b200 import metrics
    g = torch.Generator().manual_seed(B * 1000 + T)
    pred = torch.randint(lo, lo + vocab, (B, T), generator=g)
    gt = torch.randint(lo, lo + vocab, (B, T), generator=g)
    hyp, hl = metrics.caption_tokens(pred.cuda())
    ref, rl = metrics.caption_tokens(gt.cuda())
    hyps = [MO.cap_tokens(r) for r in pred.tolist()]
    refs = [MO.cap_tokens(r) for r in gt.tolist()]
    assert hl.tolist() == [len(h) for h in hyps] and rl.tolist() == [len(r) for r in refs]
    counts = metrics.bleu_counts(hyp, hl, ref, rl, 4)
    m, p, hlen, rlen = MO.bleu_counts(hyps, refs, 4)
    assert counts.tolist() == m + p + [hlen, rlen]
    for n in (1, 2, 3, 4):
        want = MO.compute_bleu(hyps, refs, n)
        assert metrics.bleu_from_counts(counts.tolist(), n) == pytest.approx(want, rel=1e-12, abs=0.0)
    # accumulation over batches == one corpus
    acc = metrics.bleu_counts(hyp[: B // 2 + 1], hl[: B // 2 + 1], ref[: B // 2 + 1], rl[: B // 2 + 1], 4)
    if B // 2 + 1 < B:
        metrics.bleu_counts(hyp[B // 2 + 1:], hl[B // 2 + 1:], ref[B // 2 + 1:], rl[B // 2 + 1:], 4, out=acc)
    assert acc.tolist() == counts.tolist()


def test_bleu_scores_end_to_end_golden():
    from hypernet_image_captioning_b200 import metrics
    c = load_case("metrics")
    got = metrics.bleu_scores(_logits(c).cuda(), c["gt_ids"].cuda())
    hyps = [c["hyp"][b, :int(c["hyp_len"][b])].tolist() for b in range(c["hyp"].shape[0])]
    refs = [c["ref"][b, :int(c["ref_len"][b])].tolist() for b in range(c["ref"].shape[0])]
    for n, v in zip((1, 2, 3, 4), got):
        assert v == pytest.approx(MO.compute_bleu(hyps, refs, n), rel=1e-12, abs=0.0)


def test_metrics_reject_cpu_tensors():
    from hypernet_image_captioning_b200 import metrics, _cabi
    with pytest.raises(_cabi.CaphnError):
        metrics.caption_tokens(torch.zeros(2, 3, dtype=torch.int64))
