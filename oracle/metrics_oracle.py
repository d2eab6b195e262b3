"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the token-level part of the reference's metric step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(hypernet_image_captioning_b200/metrics.py -> csrc/metrics.cu) never does.

  cap_tokens / cap_tokens_from_logits : utils.py:161-174 cap_to_text and :177-190 cap_to_text_gt, on token ids instead of
      words (Vocab.i2w is injective, build_vocab.py:18-24, so the filter is the same function).  PINNED: tests/golden/
      metrics.npz holds what the unmodified reference functions return (with the reference's data/vocab.pkl) for seeded
      inputs, mapped back to ids by oracle/make_golden.py.
  compute_bleu : the corpus BLEU that metric_score requests four times with max_order = 1..4 (utils.py:250-258) from
      `datasets.load_metric('bleu')`.  That metric is a third-party dependency absent from /root/reference (no version is
      pinned anywhere in the reference; `load_metric` no longer exists in the installed datasets 4.x and its script needs
      network access), so its published algorithm -- compute_bleu of tensorflow/nmt `scripts/bleu.py`, which the
      `datasets` "bleu" metric wraps unchanged -- is restated here.  PARITY UNPINNED for this function: checked only
      against hand-computed known answers in tests/test_oracle_golden.py.
"""
import collections
import math


def cap_tokens(ids, pad=0, start=1, end=2):
    """utils.py:177-190: skip <pad>/<s>, stop at </s>."""
    out = []
    for w in ids:
        w = int(w)
        if w == pad or w == start:
            continue
        if w == end:
            break
        out.append(w)
    return out


def cap_tokens_from_logits(logits, pad=0, start=1, end=2):
    """utils.py:161-174: argmax over the vocabulary axis (torch.argmax: first maximum), then the same filter."""
    import torch
    return cap_tokens(torch.argmax(logits, dim=1).tolist(), pad, start, end)


def _get_ngrams(segment, max_order):
    counts = collections.Counter()
    for order in range(1, max_order + 1):
        for i in range(0, len(segment) - order + 1):
            counts[tuple(segment[i:i + order])] += 1
    return counts


def bleu_counts(hyps, refs, max_order=4):
    """Sufficient statistics: (matches_by_order, possible_matches_by_order, translation_length, reference_length) for a
    corpus of (hypothesis, single reference) token lists."""
    matches = [0] * max_order
    possible = [0] * max_order
    ref_len = hyp_len = 0
    for hyp, ref in zip(hyps, refs):
        ref_len += len(ref)
        hyp_len += len(hyp)
        overlap = _get_ngrams(hyp, max_order) & _get_ngrams(ref, max_order)
        for ngram, c in overlap.items():
            matches[len(ngram) - 1] += c
        for order in range(1, max_order + 1):
            p = len(hyp) - order + 1
            if p > 0:
                possible[order - 1] += p
    return matches, possible, hyp_len, ref_len


def bleu_from_counts(matches, possible, hyp_len, ref_len, max_order=4):
    """compute_bleu's final arithmetic (smooth=False).  An empty hypothesis corpus (ratio 0), where the original raises
    ZeroDivisionError, gives 0.0."""
    precisions = [(matches[i] / possible[i]) if possible[i] > 0 else 0.0 for i in range(max_order)]
    geo = math.exp(sum((1.0 / max_order) * math.log(p) for p in precisions)) if min(precisions) > 0 else 0.0
    if ref_len == 0 or hyp_len == 0:
        return 0.0
    ratio = hyp_len / ref_len
    bp = 1.0 if ratio > 1.0 else math.exp(1.0 - 1.0 / ratio)
    return geo * bp


def compute_bleu(hyps, refs, max_order=4):
    m, p, hl, rl = bleu_counts(hyps, refs, max_order)
    return bleu_from_counts(m, p, hl, rl, max_order)
