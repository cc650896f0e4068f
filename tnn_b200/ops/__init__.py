"""Operator layer: hand-written sm_100a kernels (GPU) + fp32 PyTorch references (CPU oracle)."""
from . import native  # noqa: F401
