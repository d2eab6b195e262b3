"""Import shim for reference models/decoderlstm.py:11 (AttentionGru) and the DecoderGRU / DecoderRNN that hypernet.py:11
imports from it."""
from hypernet_image_captioning_b200 import AttentionGru, DecoderGRU, DecoderRNN  # noqa: F401
