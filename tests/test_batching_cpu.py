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
