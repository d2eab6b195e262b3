"""Import shim: `from hypernet_attention import HyperNet` -> B200 implementation (reference hypernet_attention.py:32)."""
from hypernet_image_captioning_b200 import HyperNetAttention as HyperNet  # noqa: F401
from hypernet_image_captioning_b200 import AttentionGru  # noqa: F401
