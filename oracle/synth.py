"""Deterministic synthetic weights / token ids (TEST INFRASTRUCTURE, see oracle/__init__.py).

The benchmark models are random-init (no network, no checkpoints; SURVEY.md
section 8d "Synthetic inputs").  Host and device must produce *bit-identical*
bf16 weights without shipping 14-93 GB of tensors, so the generator is
counter-based and uses integer arithmetic plus ONE IEEE f32 multiply chain (no
transcendental functions whose last-ulp behaviour differs between libm and CUDA):

    tensor_seed = mix64((seed * GOLDEN) ^ fnv1a64(name))
    h           = mix64(tensor_seed + (i + 1) * GOLDEN)            # element i
    s           = sum of the four 16-bit fields of h               # Irwin-Hall(4), ~normal
    value_f32   = (f32(s - 131070) * UNIT) * f32(std)              # mean 0, std `std`
    value_bf16  = round-to-nearest-even(value_f32)

The same rule is implemented in CUDA in fastllm_b200/csrc/synth.cuh (the product's
`fl_model_random_init`) and in C in oracle/synth_gen.c (fast host path).
tests/test_synth.py pins all three against each other.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

GOLDEN = 0x9E3779B97F4A7C15
MASK64 = (1 << 64) - 1
# 1 / std of the sum of four independent uniform{0..65535}: sqrt(4 * (65536^2 - 1) / 12)
UNIT = np.float32(1.0 / 37837.22719439421)
UNIT_BITS = int(np.array([UNIT], dtype=np.float32).view(np.uint32)[0])


def mix64(z: int) -> int:
    z &= MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def fnv1a64(name: str) -> int:
    h = 0xCBF29CE484222325
    for b in name.encode("utf-8"):
        h = ((h ^ b) * 0x100000001B3) & MASK64
    return h


def tensor_seed(seed: int, name: str) -> int:
    return mix64(((seed * GOLDEN) & MASK64) ^ fnv1a64(name))


def _mix64_np(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _hash_range(tseed: int, start: int, count: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + 1 + count, dtype=np.uint64)
        return _mix64_np(np.uint64(tseed) + i * np.uint64(GOLDEN))


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 bit pattern (uint16).  NaN not expected."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """f32 -> nearest bf16 value, returned as f32 (what VarBuilder(dtype=BF16) then F32 math sees)."""
    return bf16_bits_to_f32(f32_to_bf16_bits(x)).reshape(np.shape(x))


_clib = None


def _load_clib():
    global _clib
    if _clib is None:
        path = os.path.join(os.path.dirname(__file__), "_build", "liboracle_synth.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.synth_normal_bf16.argtypes = [ctypes.c_uint64, ctypes.c_float, ctypes.c_uint64,
                                              ctypes.c_uint64, ctypes.c_void_p]
            lib.synth_normal_bf16.restype = None
            _clib = lib
        else:
            _clib = False
    return _clib


def normal_bf16_bits(seed: int, name: str, numel: int, std: float = 0.02, use_c: bool = True) -> np.ndarray:
    """bf16 bit patterns (uint16) of tensor `name` under global `seed`."""
    ts = tensor_seed(seed, name)
    out = np.empty(numel, dtype=np.uint16)
    lib = _load_clib() if use_c else False
    if lib:
        lib.synth_normal_bf16(ctypes.c_uint64(ts), ctypes.c_float(std), 0, numel, out.ctypes.data)
        return out
    stdf = np.float32(std)
    step = 1 << 22
    for s0 in range(0, numel, step):
        n = min(step, numel - s0)
        h = _hash_range(ts, s0, n)
        m = np.uint64(0xFFFF)
        s = (h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))
        c = (s.astype(np.int64) - 131070).astype(np.float32)
        v = (c * UNIT) * stdf
        out[s0:s0 + n] = f32_to_bf16_bits(v)
    return out


def normal(seed: int, name: str, shape, std: float = 0.02, use_c: bool = True) -> np.ndarray:
    """bf16-rounded synthetic tensor as f32 (oracle arithmetic dtype)."""
    numel = int(np.prod(shape))
    return bf16_bits_to_f32(normal_bf16_bits(seed, name, numel, std, use_c)).reshape(shape)


def token_ids(seed: int, vocab: int, shape, name: str = "input_ids") -> np.ndarray:
    """Uniform ids in [3, vocab): avoids 0..2 (unk/bos/eos) so the "</s>" rule never fires."""
    numel = int(np.prod(shape))
    h = _hash_range(tensor_seed(seed, name), 0, numel)
    return (np.uint64(3) + h % np.uint64(vocab - 3)).astype(np.uint32).reshape(shape)
