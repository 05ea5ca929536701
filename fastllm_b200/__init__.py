"""fastllm_b200: B200-native (sm_100a) transformer forward pass behind fastllm's trait-based model API.

The product is libfastllm_b200.so (C ABI, include/fastllm_b200.h).  This package holds its sources (csrc/), the build
recipe (build.py) and the Python mirror of the reference's host interface (models.py).  Nothing here imports oracle/.
"""
from ._lib import FastllmError, FlConfig  # noqa: F401

__all__ = ["FastllmError", "FlConfig"]
