"""Import shim: `from hypernet import HyperNet` -> B200 implementation (reference hypernet.py:26, pooled variant)."""
from hypernet_image_captioning_b200 import HyperNetPooled as HyperNet  # noqa: F401
from hypernet_image_captioning_b200 import DecoderGRU  # noqa: F401
