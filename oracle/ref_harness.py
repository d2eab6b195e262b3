"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules for golden-vector generation.

The reference (zacharie12/Hypernet-image-captioning) is pure Python and depends on packages this image
lacks (pytorch_lightning, nltk, skimage, ...).  This shim installs inert stand-ins for those packages
in ``sys.modules`` and then imports the reference from ``$REFERENCE_DIR`` (default ``/root/reference``).
It is used by ``oracle/make_golden.py`` in the build container only; nothing that runs on the GPU box
imports it (``/root/reference`` does not exist there).  Nothing here is product code.

Reference entry points it exposes (all untouched reference objects):
  * ``hypernet_attention.HyperNet``            (hypernet_attention.py:32)
  * ``models.decoderlstm.AttentionGru``        (models/decoderlstm.py:11)
  * ``hypernet.HyperNet`` / ``DecoderGRU``     (hypernet.py:26, later.py:362 -- exec-injected, see load())
  * ``utils.flip_parameters_to_tensors`` / ``set_all_parameters`` (utils.py:24,44)
"""
import os
import pickle
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

import torch
from torch import nn
import torch.nn.functional as F

REFERENCE_DIR = os.environ.get("REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "hypernet_attention.py"))


class _FakeLightningModule(nn.Module):
    """Stand-in for pl.LightningModule: an nn.Module with .hparams, .log and .device."""

    def __init__(self, *a, **k):
        super().__init__()
        object.__setattr__(self, "hparams", {})

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        return torch.device("cpu")


class _TinyResNet(nn.Module):
    """Stand-in for torchvision resnet{101,152}: >=3 children and fc.in_features == 2048."""

    def __init__(self, *a, **k):
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 1)
        self.pool = nn.AdaptiveAvgPool2d(7)
        self.avg = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(2048, 10)


def install_stubs():
    """Inert stand-ins for the third-party packages the reference imports and this image lacks (no reference code is
    imported here).  Also used by tests/test_dropin_reference_scripts.py, which imports the reference *scripts* on top of
    the dropin/ shims."""
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = _FakeLightningModule
    pl.Trainer = MagicMock()
    pl.seed_everything = lambda *a, **k: None
    sys.modules["pytorch_lightning"] = pl
    for sub in ("loggers", "callbacks"):
        sys.modules[f"pytorch_lightning.{sub}"] = MagicMock()
    for name in (
        "nltk", "nltk.translate", "nltk.translate.bleu_score", "nltk.translate.meteor_score", "nltk.tokenize",
        "rouge_metric", "skimage", "skimage.io", "skimage.transform", "tldextract", "matplotlib",
        "matplotlib.pyplot", "matplotlib.image", "wandb",
    ):
        sys.modules[name] = MagicMock()
    dom = MagicMock()
    dom_tags = types.ModuleType("dominate.tags")
    dom_tags.__all__ = []
    sys.modules["dominate"] = dom
    sys.modules["dominate.tags"] = dom_tags

    import datasets  # real package; load_metric was removed in 4.x
    datasets.load_metric = lambda n, *a, **k: SimpleNamespace(name=n)
    import torchvision
    torchvision.models.resnet152 = _TinyResNet
    torchvision.models.resnet101 = _TinyResNet


_loaded = None


def load():
    """Import the reference; returns a namespace with the hot-path classes."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_DIR}")

    install_stubs()

    os.chdir(REFERENCE_DIR)  # later.py:372 opens the relative path data/vocab.pkl
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import build_vocab
    sys.modules["__main__"].Vocab = build_vocab.Vocab  # data/vocab.pkl was pickled from __main__

    import utils as ref_utils
    import models.decoderlstm as ref_dec
    import models.attention as ref_att

    # Variant A: models.decoderlstm no longer defines DecoderGRU/DecoderRNN (moved to later.py, which has
    # no imports of its own) -> exec later.py in a namespace holding what it needs, inject, then import.
    ns = {
        "torch": torch, "nn": nn, "F": F, "pickle": pickle, "np": __import__("numpy"),
        "cap_to_text_gt": ref_utils.cap_to_text_gt, "cap_to_text": ref_utils.cap_to_text,
        "sample_multinomial_topk": ref_utils.sample_multinomial_topk,
        "device": torch.device("cpu"),
    }
    for missing in ("Attention", "EncoderCNN", "BahdanauAttention"):
        ns.setdefault(missing, MagicMock())
    with open(os.path.join(REFERENCE_DIR, "later.py")) as fh:
        exec(compile(fh.read(), "later.py", "exec"), ns)
    ref_dec.DecoderGRU = ns["DecoderGRU"]
    ref_dec.DecoderRNN = ns["DecoderRNN"]

    import hypernet_attention as ref_hna
    import hypernet as ref_hn

    with open(os.path.join(REFERENCE_DIR, "data", "vocab.pkl"), "rb") as fh:
        vocab = pickle.load(fh)

    _loaded = SimpleNamespace(
        utils=ref_utils, decoderlstm=ref_dec, attention=ref_att, hypernet_attention=ref_hna, hypernet=ref_hn,
        HyperNetAttention=ref_hna.HyperNet, AttentionGru=ref_dec.AttentionGru, HyperNetPooled=ref_hn.HyperNet,
        DecoderGRU=ref_dec.DecoderGRU, vocab=vocab,
    )
    return _loaded
