"""Token-level half of the reference's metric step on the device (kernels in csrc/metrics.cu).

The reference converts every predicted and ground-truth caption to words with one ``.item()`` per token
(utils.py:161-190, called from metric_score utils.py:229-262 inside every training_step, cc_train_hypernet.py:154).
Here the argmax, the <pad>/<s>/</s> filter and the BLEU n-gram statistics run on the device for the whole batch; the
host receives one small tensor.  Word strings (needed by METEOR / ROUGE / CIDEr, which stay CPU string code and are out
of scope) come from ``tokens_to_text`` after that single copy.
"""
import math
from typing import List, Sequence, Tuple

import torch

from . import _cabi, ops


def caption_tokens(pred: torch.Tensor, pad: int = 0, start: int = 1, end: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    """cap_to_text (utils.py:161-174) / cap_to_text_gt (:177-190) on token ids for a whole batch.

    ``pred``: logits or probabilities ``[B, T, V]`` (argmax on the device, lowest index on ties like ``torch.argmax``)
    or token ids ``[B, T]`` int64 (ground truth).  Returns ``(tokens [B, T] int64, lengths [B] int32)``: row b holds the
    kept tokens -- neither ``pad`` nor ``start``, up to the first ``end`` -- followed by ``pad``."""
    if pred.dim() == 3:
        B, T, V = pred.shape
        x = pred.reshape(B * T, V)
        if x.stride(1) != 1:
            x = x.contiguous()
        _, ids = ops.softmax_argmax(x.float() if x.dtype != torch.float32 else x, want_probs=False)
        ids = ids.view(B, T)
    else:
        ids = pred.to(torch.int64).contiguous()
        B, T = ids.shape
    ops._chk(ids, torch.int64)
    out = torch.empty(B, T, device=ids.device, dtype=torch.int64)
    lens = torch.empty(B, device=ids.device, dtype=torch.int32)
    _cabi.call("caphn_caption_compact", ids.data_ptr(), ids.stride(0), B, T, pad, start, end, out.data_ptr(),
               lens.data_ptr(), ops._stream())
    return out, lens


def bleu_counts(hyp: torch.Tensor, hyp_len: torch.Tensor, ref: torch.Tensor, ref_len: torch.Tensor,
                max_order: int = 4, out: torch.Tensor = None) -> torch.Tensor:
    """Corpus-BLEU sufficient statistics of B (hypothesis, reference) pairs of compacted captions, on the device:
    int64 ``[2*max_order + 2]`` = clipped n-gram matches per order, candidate n-grams per order, hypothesis length,
    reference length.  Pass ``out`` to keep accumulating over batches (an epoch-level corpus BLEU)."""
    ops._chk(hyp, torch.int64), ops._chk(ref, torch.int64)
    hyp, ref = hyp.contiguous(), ref.contiguous()
    hyp_len, ref_len = hyp_len.to(torch.int32).contiguous(), ref_len.to(torch.int32).contiguous()
    B = hyp.shape[0]
    assert ref.shape[0] == B and hyp_len.numel() == B and ref_len.numel() == B
    if out is None:
        out = torch.zeros(2 * max_order + 2, device=hyp.device, dtype=torch.int64)
    _cabi.call("caphn_bleu_counts", hyp.data_ptr(), hyp_len.data_ptr(), hyp.shape[1], ref.data_ptr(), ref_len.data_ptr(),
               ref.shape[1], B, max_order, out.data_ptr(), ops._stream())
    return out


def bleu_from_counts(counts: Sequence[int], order: int, max_order: int = 4) -> float:
    """BLEU-``order`` from the statistics of ``bleu_counts`` (host arithmetic on ten integers): geometric mean of the
    first ``order`` clipped precisions times the brevity penalty -- `datasets` "bleu" / tensorflow-nmt compute_bleu with
    ``max_order=order, smooth=False``, which is what metric_score asks for (utils.py:250-258)."""
    c = [int(v) for v in counts]
    matches, possible, hl, rl = c[:max_order], c[max_order:2 * max_order], c[2 * max_order], c[2 * max_order + 1]
    prec = [(matches[i] / possible[i]) if possible[i] > 0 else 0.0 for i in range(order)]
    geo = math.exp(sum(math.log(p) / order for p in prec)) if min(prec) > 0 else 0.0
    if hl == 0 or rl == 0:
        return 0.0
    ratio = hl / rl
    return geo * (1.0 if ratio > 1.0 else math.exp(1.0 - 1.0 / ratio))


def bleu_scores(pred: torch.Tensor, gt_caps: torch.Tensor, pad: int = 0, start: int = 1, end: int = 2) -> List[float]:
    """[BLEU-1, BLEU-2, BLEU-3, BLEU-4] of a batch: the first four entries of metric_score's output
    (utils.py:229-262 with the 'bleu' metric), from one device->host copy of ten integers."""
    hyp, hl = caption_tokens(pred, pad, start, end)
    ref, rl = caption_tokens(gt_caps, pad, start, end)
    counts = bleu_counts(hyp, hl, ref, rl, 4).tolist()
    return [bleu_from_counts(counts, n) for n in (1, 2, 3, 4)]


def tokens_to_text(tokens: torch.Tensor, lengths: torch.Tensor, vocab, tokenized: bool = False):
    """Words of a batch of compacted captions (``vocab.i2w`` as in build_vocab.py:18-24): one device->host copy for the
    batch instead of one ``.item()`` per token."""
    toks, lens = tokens.cpu().tolist(), lengths.cpu().tolist()
    sents = [[vocab.i2w[w] for w in row[:n]] for row, n in zip(toks, lens)]
    return sents if tokenized else [" ".join(s) for s in sents]
