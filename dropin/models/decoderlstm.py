"""Import shim for reference models/decoderlstm.py:11 (AttentionGru) and the DecoderGRU that hypernet.py:10 imports from it."""
from hypernet_image_captioning_b200 import AttentionGru, DecoderGRU  # noqa: F401
