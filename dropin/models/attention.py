"""Import shim for reference models/attention.py:5 (BahdanauAttention: parameter holder; the math runs in the recurrence)."""
from hypernet_image_captioning_b200 import BahdanauAttention  # noqa: F401
