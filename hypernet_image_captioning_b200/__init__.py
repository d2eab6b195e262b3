"""caption-hn-b200: B200-native (sm_100a) hot path of Caption-HN -- hypernetwork -> generated GRU weights -> decoder.

Public API mirrors the reference modules (see modules.py / modules_attention.py); kernels live in csrc/ behind the
C-ABI declared in include/caphn_b200.h.  Importing the package does not need a GPU; calling any op does.
"""
from . import _cabi, ops, functional, graphs, metrics  # noqa: F401
from .functional import cross_entropy, linear, hypernet_theta  # noqa: F401
from .modules import DecoderGRU, DecoderRNN, HyperNetPooled, PooledFeatureEncoder  # noqa: F401
from .modules_attention import AttentionGru, BahdanauAttention, HyperNetAttention, SpatialFeatureEncoder  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .style import DomainEmbedding  # noqa: F401

__all__ = ["FusedAdam", "DomainEmbedding", "AttentionGru", "BahdanauAttention", "HyperNetAttention", "SpatialFeatureEncoder",
           "DecoderGRU", "DecoderRNN", "HyperNetPooled", "PooledFeatureEncoder", "cross_entropy", "linear", "hypernet_theta"]
