"""CPU restatement of the sampling step of the reference's generate loop.  TEST INFRASTRUCTURE ONLY.

Reference call sites (paths under /root/reference/src):
    LogitsProcessor::new(Default::default(), Some(temperature as f64), None)     models/mod.rs:157-158, 373-374
    logits_processor.sample(&last_logits)                                         models/mod.rs:308-310, 425-428

The arithmetic lives in un-vendored crates (Cargo.toml:19-21, no lockfile): candle-transformers 0.8.x
`generation::LogitsProcessor`, candle-nn `ops::softmax_last_dim`, and -- through candle -- rand 0.8.5
(`StdRng` = ChaCha12, `SeedableRng::seed_from_u64`, `distributions::WeightedIndex<f32>`, `Uniform<f32>`).
Their published algorithms are restated below, one function per piece:

    seed_from_u64      rand_core 0.6 `SeedableRng::seed_from_u64`: PCG32 (MUL 6364136223846793005, INC 11634580027462260723),
                       one output word per 4 seed bytes, little endian
    ChaCha12           rand_chacha 0.3: djb layout -- 4 constants, 8 key words, 64-bit block counter (words 12-13),
                       64-bit stream id (words 14-15, 0), 12 rounds; `next_u32` hands out the key-stream words in order
    Uniform<f32>       rand 0.8.5 `UniformFloat::new(low, high)`: scale = high - low, decreased one ulp at a time while
                       scale * (1 - 2^-23) + low >= high; sample = ((u32 >> 9 | 0x3f800000 as f32) - 1.0) * scale + low
    WeightedIndex<f32> rand 0.8.5: cumulative weights accumulated left to right in f32 (the last weight only enters the
                       total), sample = partition_point(|w| w <= chosen)
    softmax_last_dim   candle-nn CPU kernel: max by fold, d = exp(s - max), sum_exp = sequential f32 sum, d /= sum_exp
    logits / T         candle `Tensor / f64` = affine(1/T, 0): v * (1/T as f32) + 0.0
    sample_argmax      iter().enumerate().max_by(|u, v| u.total_cmp(v)): LAST index among equal maxima, IEEE total order
                       (so a positive NaN beats +inf)

PARITY STATUS: pinned for the generator (rand's own value-stability vector for StdRng and the published ChaCha
key-stream vectors, tests/test_sampling_cpu.py); **unpinned for the float pipeline** (softmax -> WeightedIndex), which has
no vector anywhere in the reference and cannot be run here (no Rust toolchain).  `exp` is the C library's `expf` (what
Rust's `f32::exp` links on linux-gnu), called through ctypes; the reduce kernels are candle's scalar ones (`vec_reduce_max`
/ `vec_reduce_sum` without the avx/neon target features -- the reference sets no RUSTFLAGS / target-cpu).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import struct

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.expf.restype, _libm.expf.argtypes = ctypes.c_float, [ctypes.c_float]
_expf = np.frompyfunc(lambda v: _libm.expf(float(v)), 1, 1)

F32 = np.float32
_M32 = 0xFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF


def seed_from_u64(state: int) -> bytes:
    """rand_core 0.6 SeedableRng::seed_from_u64 for a 32-byte seed."""
    out = b""
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & _M64
        xorshifted = ((((state >> 18) ^ state) >> 27)) & _M32
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & _M32
        out += struct.pack("<I", x)
    return out


def _rotl(v: int, n: int) -> int:
    return ((v << n) | (v >> (32 - n))) & _M32


def chacha_block(key_words, counter: int, stream: int = 0, rounds: int = 12):
    """One 64-byte ChaCha block as 16 u32 words (djb variant: 64-bit counter, 64-bit stream id)."""
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *key_words,
            counter & _M32, (counter >> 32) & _M32, stream & _M32, (stream >> 32) & _M32]
    x = list(init)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & _M32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & _M32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & _M32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & _M32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & _M32 for a, b in zip(x, init)]


class StdRng:
    """rand 0.8 StdRng (ChaCha12Rng behind a BlockRng): the key-stream words, in order, from block counter 0."""

    def __init__(self, seed: bytes, rounds: int = 12):
        assert len(seed) == 32
        self.key = list(struct.unpack("<8I", seed))
        self.rounds = rounds
        self.counter = 0
        self.buf: list = []

    @classmethod
    def seed_from_u64(cls, state: int) -> "StdRng":
        return cls(seed_from_u64(state))

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha_block(self.key, self.counter, 0, self.rounds)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


def uniform_f32_scale(low: F32, high: F32) -> F32:
    """rand 0.8.5 UniformFloat::<f32>::new(low, high) -> scale."""
    max_rand = F32(1.0) - F32(2.0 ** -23)
    scale = F32(high - low)
    while F32(F32(scale * max_rand) + low) >= high:
        scale = np.frombuffer(struct.pack("<I", struct.unpack("<I", struct.pack("<f", scale))[0] - 1), dtype=F32)[0]
    return scale


def uniform_f32_sample(rng: StdRng, low: F32, scale: F32) -> F32:
    value1_2 = np.frombuffer(struct.pack("<I", (rng.next_u32() >> 9) | 0x3F800000), dtype=F32)[0]
    return F32(F32(F32(value1_2 - F32(1.0)) * scale) + low)


def weighted_index_sample(rng: StdRng, weights: np.ndarray) -> int:
    """WeightedIndex::<f32>::new(weights)?.sample(rng).  Raises ValueError where rand returns WeightedError."""
    w = np.asarray(weights, dtype=F32).reshape(-1)
    if w.size == 0:
        raise ValueError("NoItem")
    if not bool(np.all(w >= 0)):          # a NaN fails `w >= zero` too
        raise ValueError("InvalidWeight")
    # cumulative[i] = w[0] + ... + w[i] accumulated left to right in f32 (np.cumsum on f32 is a sequential f32 scan)
    cum = np.cumsum(w, dtype=F32)
    total = cum[-1]
    if total == 0:
        raise ValueError("AllWeightsZero")
    chosen = uniform_f32_sample(rng, F32(0), uniform_f32_scale(F32(0), total))
    # partition_point(|c| c <= chosen) over cumulative[..n-1]
    return int(np.searchsorted(cum[:-1], chosen, side="right"))


def softmax_last_dim(x: np.ndarray) -> np.ndarray:
    """candle-nn softmax_last_dim CPU kernel on one row: sequential f32 sum of the exponentials."""
    x = np.asarray(x, dtype=F32).reshape(-1)
    with np.errstate(invalid="ignore"):
        d = _expf((x - np.fmax.reduce(x)).astype(F32)).astype(F32)     # f32::max ignores a NaN operand, like np.fmax
    sum_exp = np.cumsum(d, dtype=F32)[-1]
    return (d / sum_exp).astype(F32)


def _total_order_key(v: np.ndarray) -> np.ndarray:
    bits = v.view(np.int32).astype(np.int64)
    return np.where(bits < 0, bits ^ 0x7FFFFFFF, bits)


def sample_argmax(logits: np.ndarray) -> int:
    v = np.ascontiguousarray(logits, dtype=F32).reshape(-1)
    k = _total_order_key(v)
    return int(np.flatnonzero(k == k.max())[-1])


class LogitsProcessor:
    """candle-transformers 0.8 LogitsProcessor::new(seed, temperature, None): ArgMax when temperature is None or < 1e-7,
    else Sampling::All { temperature }."""

    def __init__(self, seed: int = 0, temperature: float | None = None):
        self.rng = StdRng.seed_from_u64(seed)
        self.temperature = None if temperature is None or temperature < 1e-7 else float(temperature)

    def sample(self, logits: np.ndarray) -> int:
        v = np.asarray(logits, dtype=F32).reshape(-1)
        if self.temperature is None:
            return sample_argmax(v)
        inv_t = F32(1.0 / self.temperature)
        prs = softmax_last_dim(F32(v * inv_t) + F32(0.0))
        return weighted_index_sample(self.rng, prs)
