"""Request batching at the API boundary: host logic, tested on CPU against a stand-in model.

The batcher is exact by construction -- only requests of equal token length share a [b, t] call -- so the property under test is
generate_batch(prompts) == [generate(p) for p in prompts] and embed_many(sentences)[i] == embed(sentences[i]), in request order,
for greedy and temperature sampling, with per-sequence EOS."""
import zlib

import numpy as np
import pytest

from fastllm_b200 import models

VOCAB = 97


class _Cache:
    def __init__(self):
        self.hist = None


class _HistoryModel:
    """Logits of a row are a pure function of that row's token history (what a causal LM is), so batching must not change them."""

    def __init__(self, eos_bias=0.0, three_d=False):
        self.calls, self.eos_bias, self.three_d = [], eos_bias, three_d

    def initialize_cache(self, *a):
        return _Cache()

    def forward(self, ids, pos, cache):
        ids = np.asarray(ids)
        self.calls.append((ids.shape, pos))
        if pos == 0:
            cache.hist = [list(map(int, r)) for r in ids]
        else:
            assert ids.shape[1] == 1 and len(cache.hist) == ids.shape[0] and pos == len(cache.hist[0])
            for h, r in zip(cache.hist, ids):
                h.append(int(r[0]))
        rows = []
        for h in cache.hist:
            rs = np.random.RandomState(zlib.crc32(np.asarray(h, dtype=np.uint32).tobytes()))
            row = (rs.standard_normal(VOCAB) * 2).astype(np.float32)
            row[2] += self.eos_bias
            rows.append(row)
        out = np.stack(rows)
        return out[:, None, :] if self.three_d else out


def _prompts(seed, lengths):
    rs = np.random.RandomState(seed)
    return [list(map(int, rs.randint(3, VOCAB, size=n))) for n in lengths]


@pytest.mark.parametrize("temperature", [0.0, 0.9])
@pytest.mark.parametrize("three_d", [False, True])
def test_generate_batch_equals_per_request(temperature, three_d):
    prompts = _prompts(1, [5, 9, 5, 5, 3, 9, 7, 5])
    m = _HistoryModel(eos_bias=2.5, three_d=three_d)          # "</s>" (id 2) is likely enough that several sequences stop early
    single = [models.Model(m, None, eos_token_id=2).generate(p, 12, temperature=temperature) for p in prompts]
    assert len({len(s) for s in single}) > 1, "the case must include sequences that hit EOS at different steps"
    m2 = _HistoryModel(eos_bias=2.5, three_d=three_d)
    batched = models.Model(m2, None, eos_token_id=2).generate_batch(prompts, 12, temperature=temperature)
    assert batched == single
    # one prefill per length group (5: four prompts, 9: two, 3: one, 7: one), batches never mix lengths
    prefill_shapes = [shape for shape, pos in m2.calls if pos == 0]
    assert prefill_shapes == [(4, 5), (2, 9), (1, 3), (1, 7)]
    assert all(shape[1] == 1 for shape, pos in m2.calls if pos > 0)


def test_generate_batch_respects_max_batch_and_order():
    prompts = _prompts(2, [4] * 7)
    m = _HistoryModel()
    single = [models.Model(m, None, eos_token_id=None).generate(p, 3) for p in prompts]
    m2 = _HistoryModel()
    assert models.Model(m2, None, eos_token_id=None).generate_batch(prompts, 3, max_batch=3) == single
    assert [shape for shape, pos in m2.calls if pos == 0] == [(3, 4), (3, 4), (1, 4)]


def test_generate_batch_empty_and_zero_tokens():
    m = _HistoryModel()
    assert models.Model(m, None).generate_batch([], 4) == []
    assert models.Model(m, None).generate_batch(_prompts(3, [2, 2]), 0) == [[], []]


def test_embed_many_groups_equal_lengths_and_scatters_back():
    calls = []

    def fake_embed_ids(ids, mask=None):
        ids = np.asarray(ids)
        calls.append(ids.shape)
        assert mask is None                                    # no padding is ever introduced
        v = np.stack([np.cos(np.arange(8) * (1 + int(r.sum()))) for r in ids]).astype(np.float32)
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    m = models.MiniLMModel.__new__(models.MiniLMModel)        # host logic only: no device model behind it
    m.config = models.BertConfig(hidden_size=8)
    m.embed_ids = fake_embed_ids
    sentences = _prompts(4, [6, 3, 6, 11, 3, 6, 6])
    got = m.embed_many(sentences, max_batch=3)
    want = np.concatenate([fake_embed_ids(np.asarray([s])) for s in sentences])
    assert np.array_equal(got, want)
    assert calls[:4] == [(3, 6), (1, 6), (2, 3), (1, 11)]
    assert m.embed_many([]).shape == (0, 8)


class _SlotCache:
    """Stand-in for DeviceCache's slot calls: logits of a row are a pure function of that slot's token history."""

    def __init__(self, nslots):
        self.hist = [None] * nslots
        self.calls = []

    def forward_slots(self, slots, ids, rope_offsets):
        ids = np.asarray(ids)
        self.calls.append((tuple(int(s) for s in slots), ids.shape, tuple(int(r) for r in rope_offsets)))
        assert len(set(slots)) == len(slots) == ids.shape[0]
        rows = []
        for s, r, ro in zip(slots, ids, rope_offsets):
            if self.hist[s] is None:
                assert ro == 0
                self.hist[s] = []
            else:
                assert ids.shape[1] == 1 and ro == len(self.hist[s])          # Llama rule: the caller's position
            self.hist[s] += [int(x) for x in r]
            rs = np.random.RandomState(zlib.crc32(np.asarray(self.hist[s], dtype=np.uint32).tobytes()))
            row = (rs.standard_normal(VOCAB) * 2).astype(np.float32)
            row[2] += 2.5
            rows.append(row)
        return np.stack(rows)

    def slot_reset(self, slot):
        assert self.hist[slot] is not None
        self.hist[slot] = None


class _Llama:
    arch = "llama"


@pytest.mark.parametrize("temperature", [0.0, 0.9])
def test_continuous_batcher_equals_per_request_generate(temperature):
    """Requests of different lengths share decode steps, slots are reused as requests finish (8 requests over 3 slots), and every
    request gets exactly the tokens Model.generate gives it alone."""
    prompts = _prompts(5, [5, 9, 2, 7, 3, 11, 4, 6])
    m = _HistoryModel(eos_bias=2.5)
    single = [models.Model(m, None, eos_token_id=2).generate(p, 10, temperature=temperature) for p in prompts]
    assert len({len(s) for s in single}) > 1
    cache = _SlotCache(3)
    cb = models.ContinuousBatcher(_Llama(), max_batch=3, eos_token_id=2, cache=cache)
    assert cb.generate(prompts, 10, temperature=temperature) == single
    prefills = [c for c in cache.calls if c[1][1] > 1 or c[2] == (0,)]
    assert len(prefills) == len(prompts) and all(len(c[0]) == 1 for c in prefills)       # prompts are prefilled alone
    decodes = [c for c in cache.calls if c not in prefills]
    assert max(len(c[0]) for c in decodes) == 3 and all(c[1][1] == 1 for c in decodes)   # ragged steps fill the slots
    assert all(h is None for h in cache.hist)                                            # every slot was handed back
    assert cb.steps < sum(len(s) for s in single)                                        # steps were shared


def test_continuous_batcher_position_rule_of_the_mistral_adapter_and_edge_cases():
    class _Mistral:
        arch = "mistral"

    class _C(_SlotCache):
        def forward_slots(self, slots, ids, rope_offsets):
            self.calls.append((tuple(slots), np.asarray(ids).shape, tuple(int(r) for r in rope_offsets)))
            return np.tile(np.arange(VOCAB, dtype=np.float32), (len(slots), 1))           # arg-max = VOCAB - 1, never EOS

        def slot_reset(self, slot):
            pass

    cache = _C(2)
    out = models.ContinuousBatcher(_Mistral(), max_batch=2, eos_token_id=2, cache=cache).generate([[5, 6, 7], [8]], 3)
    assert out == [[VOCAB - 1] * 3, [VOCAB - 1] * 3]
    # mistral.rs:234 / qwen.rs:143: the offset grows by ONE per call, whatever the prompt length: 0 at the prefill, then 1, 2
    assert [c[2] for c in cache.calls] == [(0,), (0,), (1, 1), (2, 2)]
    assert models.ContinuousBatcher(_Mistral(), 2, cache=_C(2)).generate([], 4) == []
    assert models.ContinuousBatcher(_Mistral(), 2, cache=_C(2)).generate([[3, 4]], 0) == [[]]
