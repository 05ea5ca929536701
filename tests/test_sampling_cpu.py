"""Sampling step of the generate loop: candle's LogitsProcessor over rand 0.8's StdRng.

CPU tests: the generator is pinned by published known-answer vectors (the ChaCha key stream for the all-zero key, and rand
0.8's own value-stability vector for StdRng); the product's sampler (host code inside libfastllm_b200.so, reached through the
C ABI -- no GPU involved) must then reproduce the oracle's token stream bit for bit."""
import ctypes as C
import struct

import numpy as np
import pytest

from fastllm_b200 import _lib, models
from oracle import sampling as osamp


# ---- the generator, against published vectors ---------------------------------------------------------------------------
CHACHA_ZERO_KEY = {   # first 32 key-stream bytes, all-zero key / nonce / counter (the ChaCha test-vector set, TC1)
    20: "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7",
    12: "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f",
    8: "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e",
}


@pytest.mark.parametrize("rounds", [20, 12, 8])
def test_oracle_chacha_block_known_answer(rounds):
    words = osamp.chacha_block([0] * 8, 0, 0, rounds)
    assert struct.pack("<16I", *words)[:32].hex() == CHACHA_ZERO_KEY[rounds]


def test_oracle_stdrng_value_stability_vector():
    """rand 0.8 `test_stdrng_construction`: the seed, next_u64 of it, and next_u64 of StdRng::from_rng(it)."""
    seed = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)
    rng0 = osamp.StdRng(seed)
    x0 = rng0.next_u64()
    rng1 = osamp.StdRng(b"".join(struct.pack("<I", rng0.next_u32()) for _ in range(8)))     # from_rng: fill_bytes of a 32-byte seed
    assert [x0, rng1.next_u64()] == [10719222850664546238, 14064965282130556830]


def test_oracle_seed_from_u64_is_pcg32():
    # PCG32 (XSH-RR) reference stream for state 0 / this increment, computed independently with Python integers
    state, words = 0, []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) % (1 << 64)
        xs, rot = ((state >> 18 ^ state) >> 27) % (1 << 32), state >> 59
        words.append((xs >> rot | xs << (-rot & 31)) % (1 << 32))
    assert osamp.seed_from_u64(0) == struct.pack("<8I", *words)
    assert len({osamp.seed_from_u64(s) for s in (0, 1, 2, 3, 4, 8, 16, 2**64 - 1)}) == 8


@pytest.mark.parametrize("seed", [0, 1, 299792458, 2**64 - 1])
def test_product_generator_stream_equals_oracle(seed):
    lp = models.LogitsProcessor(seed, 1.0)
    ref = osamp.StdRng.seed_from_u64(seed)
    assert [lp.next_u32() for _ in range(200)] == [ref.next_u32() for _ in range(200)]     # 12.5 blocks


# ---- arg-max (temperature None / < 1e-7) ---------------------------------------------------------------------------------
ARGMAX_CASES = [
    ([1.0, 3.0, 3.0, 2.0, 3.0, 0.0], 4),                 # max_by keeps the LAST maximum
    ([5.0], 0),
    ([0.0, -0.0], 0), ([-0.0, 0.0], 1),                   # total_cmp: -0 < +0
    ([1.0, float("inf"), float("nan"), 2.0], 2),          # a positive NaN is above +inf in the total order
    ([-float("inf")] * 3, 2),
]


@pytest.mark.parametrize("vals,want", ARGMAX_CASES)
def test_argmax_total_order(vals, want):
    v = np.array(vals, dtype=np.float32)
    assert osamp.sample_argmax(v) == want
    assert models.LogitsProcessor(0, None).sample(v) == want
    assert models.LogitsProcessor(0, 0.0).sample(v) == want
    assert models.LogitsProcessor(0, 9e-8).sample(v) == want       # below the 1e-7 threshold
    assert models.sample_argmax(v) == want


def test_negative_nan_is_below_everything():
    v = np.array([-1.0, 0.0], dtype=np.float32)
    v.view(np.uint32)[1] = 0xFFC00000                              # -NaN
    assert osamp.sample_argmax(v) == 0 == models.sample_argmax(v)


# ---- temperature sampling -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vocab,temperature,seed", [(7, 1.0, 0), (1000, 0.7, 0), (32000, 1.0, 0), (32000, 1.5, 3), (152064, 0.3, 0),
                                                     (4096, 1e-7, 1)])
def test_product_token_stream_equals_oracle(vocab, temperature, seed):
    """Same seed, same logits every step => identical token ids, draw after draw (the generator state carries over)."""
    rs = np.random.RandomState(vocab)
    want_lp, got_lp = osamp.LogitsProcessor(seed, temperature), models.LogitsProcessor(seed, temperature)
    for _ in range(12):
        logits = (rs.standard_normal(vocab) * 1.3).astype(np.float32)
        assert got_lp.sample(logits) == want_lp.sample(logits)


def test_sampling_follows_the_distribution():
    """Size-independent property: empirical frequencies over many draws match softmax(logits / T)."""
    logits = np.array([2.0, 1.0, 0.0, -1.0, 3.0, 3.0], dtype=np.float32)
    t = 0.8
    p = np.exp(logits / t - (logits / t).max())
    p /= p.sum()
    lp = models.LogitsProcessor(0, t)
    n = 40000
    counts = np.bincount([lp.sample(logits) for _ in range(n)], minlength=logits.size)
    assert np.abs(counts / n - p).max() < 4 * np.sqrt(p.max() * (1 - p.max()) / n) + 1e-3


def test_one_hot_distribution_never_picks_a_zero_weight():
    logits = np.full(5000, -1e4, dtype=np.float32)      # exp underflows to exactly 0 for everything but one entry
    logits[1234] = 0.0
    lp = models.LogitsProcessor(0, 1.0)
    assert {lp.sample(logits) for _ in range(64)} == {1234}
    assert osamp.LogitsProcessor(0, 1.0).sample(logits) == 1234


def test_uniform_scale_shrinks_below_high():
    """UniformFloat::new(0, total): the largest possible sample must stay below `total`."""
    for total in [np.float32(1.0), np.float32(0.99999994), np.float32(1.0000001), np.float32(3.0)]:
        scale = osamp.uniform_f32_scale(np.float32(0), total)
        assert np.float32(scale * (np.float32(1) - np.float32(2.0 ** -23))) < total


@pytest.mark.parametrize("bad", [[0.0, float("nan"), 1.0], [float("inf"), 0.0], [-float("inf")] * 3])
def test_failure_points_match(bad):
    """Where rand returns WeightedError (NaN probabilities), the reference's `?` turns it into an Err: both sides must fail."""
    v = np.array(bad, dtype=np.float32)
    with pytest.raises(ValueError):
        osamp.LogitsProcessor(0, 1.0).sample(v)
    with pytest.raises(_lib.FastllmError) as ei:
        models.LogitsProcessor(0, 1.0).sample(v)
    assert ei.value.code == -1 and "Weight" in str(ei.value)


def test_null_and_empty_arguments_are_errors():
    lib = _lib.load()
    assert lib.fl_sampler_create(0, 1.0, None) != 0
    assert lib.fl_sampler_create(0, float("nan"), C.byref(C.c_void_p())) != 0
    lp = models.LogitsProcessor(0, 1.0)
    with pytest.raises(_lib.FastllmError):
        lp.sample(np.zeros(0, dtype=np.float32))


def test_cpp_host_mirror_logits_processor(tmp_path):
    """host/fastllm_host.hpp's LogitsProcessor (the compiled mirror) draws the same ids as the oracle; no GPU involved."""
    import json
    import os
    import subprocess
    import __graft_entry__ as ge
    ge._build_host_mirror()
    exe = os.path.join(ge.ROOT, "host", "_build", "host_selftest")
    rs = np.random.RandomState(11)
    rows = (rs.standard_normal((10, 2048)) * 1.7).astype(np.float32)
    path = tmp_path / "logits.bin"
    rows.tofile(path)
    for temperature in (0.8, 0.0, -1.0):
        out = subprocess.run([exe, "--sampler", "0", str(temperature), "2048", str(path)], capture_output=True, text=True, check=True)
        want_lp = osamp.LogitsProcessor(0, None if temperature < 0 else temperature)
        assert json.loads(out.stdout) == [want_lp.sample(r) for r in rows]


class _ScriptedAdapter:
    """A stand-in model: forward returns a fixed logits row per call (the generate loop is host logic and must not care)."""

    def __init__(self, rows, three_d):
        self.rows, self.calls, self.three_d = rows, [], three_d

    def initialize_cache(self, *a):
        return object()

    def forward(self, ids, pos, cache):
        self.calls.append((np.asarray(ids).copy(), pos))
        row = self.rows[len(self.calls) - 1][None]
        return row[:, None, :] if self.three_d else row


@pytest.mark.parametrize("three_d", [False, True])
def test_generate_loop_with_temperature_matches_oracle(three_d):
    rs = np.random.RandomState(5)
    rows = [(rs.standard_normal(300) * 2).astype(np.float32) for _ in range(20)]
    rows[6][:] = -50.0
    rows[6][2] = 50.0                                              # step 6 samples "</s>" (id 2) with probability 1
    want_lp = osamp.LogitsProcessor(0, float(np.float32(0.9)))
    want = []
    for r in rows:
        tok = want_lp.sample(r)
        if tok == 2:
            break
        want.append(tok)
    ad = _ScriptedAdapter(rows, three_d)
    got = models.Model(ad, None, eos_token_id=2).generate([5, 6, 7], 19, temperature=0.9)
    assert got == want and len(got) == 6                           # EOS breaks before emitting
    assert [p for _, p in ad.calls] == [0, 3, 4, 5, 6, 7, 8]       # pos: 0, then prompt length + i
    assert ad.calls[0][0].shape == (1, 3) and all(c.shape == (1, 1) for c, _ in ad.calls[1:])


def test_argmax_rows_equals_row_by_row():
    """fl_argmax_rows (a batch of greedy requests, threaded over rows above 1 Mi elements) == the single-row rule, ties and all."""
    rs = np.random.RandomState(9)
    for b, v in [(1, 7), (5, 1000), (64, 32000)]:
        x = rs.standard_normal((b, v)).astype(np.float32)
        x[b // 2, [0, v // 3, v - 1]] = 9.0                    # a three-way tie: the LAST index wins
        x[0, v // 2] = np.inf
        got = models.sample_argmax_rows(x)
        assert got.dtype == np.uint32 and got.shape == (b,)
        assert list(got) == [osamp.sample_argmax(r) for r in x]
    assert int(models.sample_argmax_rows(x)[b // 2]) == v - 1
    assert models.sample_argmax_rows(x[:, None, :]).shape == (b,)       # [b, 1, V] logits of the Mistral / Qwen2 adapters


# Frozen vectors of THIS repo's restatement (generated by the oracle when it was written, after the generator had been pinned by the
# published vectors above): they keep the oracle and the product from drifting together.  Not reference outputs -- the float
# pipeline stays "parity unpinned" (oracle/sampling.py header).
FROZEN_STREAMS = {0: [3442241407, 3140108210, 2384947579, 3321986196, 3476097558, 111001858],
                  42: [572990626, 2261546851, 1068323197, 2330987027]}
FROZEN_TOKENS = {(1000, 0.7, 0): [780, 767, 419, 739, 822, 27, 634, 625, 834, 317, 154, 777],
                 (32000, 1.0, 0): [25538, 23141, 17694, 24594, 25805, 820, 21744, 19024, 27848, 8498, 3989, 24711],
                 (7, 1.0, 0): [3, 4, 6, 2, 4, 1, 5, 1, 4, 0, 0, 4]}


def test_frozen_generator_streams():
    for seed, want in FROZEN_STREAMS.items():
        ref, lp = osamp.StdRng.seed_from_u64(seed), models.LogitsProcessor(seed, 1.0)
        assert [ref.next_u32() for _ in want] == want == [lp.next_u32() for _ in want]


@pytest.mark.parametrize("key", sorted(FROZEN_TOKENS))
def test_frozen_token_streams(key):
    vocab, temperature, seed = key
    for make in (osamp.LogitsProcessor, models.LogitsProcessor):
        rs, lp = np.random.RandomState(vocab), make(seed, temperature)
        assert [lp.sample((rs.standard_normal(vocab) * 1.3).astype(np.float32)) for _ in range(12)] == FROZEN_TOKENS[key]
