"""Mixtral sparse-MoE block restated in numpy f32 (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference does not wire Mixtral at all (src/models/model_registry.rs:169-182 has no key that
"MixtralForCausalLM" contains; src/models/mistral.rs:244-246 accepts only MistralForCausalLM), so this follows
candle-transformers 0.8.x models::mixtral::SparseMoeBlock::forward (un-vendored):

    router_logits = x . W_gate^T                      [n, E]
    routing       = softmax_last_dim(router_logits)   over ALL experts, f32
    per row: indices sorted by descending routing weight (stable => ties keep the LOWER expert index first),
             take top_k, renormalise the chosen weights by their sum (f32 on the host)
    per expert e: rows = index_select(x, rows_e); y = w2(silu(w1 rows) * w3 rows); y *= weight; index_add into out
"""
from __future__ import annotations

import numpy as np

from . import candle_ops as ops

F32 = np.float32


def route_top_k(router_logits: np.ndarray, top_k: int = 2):
    """-> (expert index [n, top_k] int64, renormalised weight [n, top_k] f32)."""
    probs = ops.softmax_last_dim(router_logits.astype(F32))
    n, _ = probs.shape
    idx = np.empty((n, top_k), dtype=np.int64)
    wts = np.empty((n, top_k), dtype=F32)
    for r in range(n):
        order = np.argsort(-probs[r], kind="stable")[:top_k]      # sort_by(|i,j| rw[j].total_cmp(rw[i])), stable
        chosen = probs[r][order]
        s = F32(0.0)
        for c in chosen:                                          # sum::<f32>() in index order
            s = F32(s + c)
        idx[r] = order
        wts[r] = chosen / s
    return idx, wts


ROUTING_LOG = None      # tests set this to a list to record the router's picks (checked against fl_cache_moe_routing)


def sparse_moe_block(x: np.ndarray, w_gate: np.ndarray, experts, top_k: int = 2) -> np.ndarray:
    """x [b, t, H]; experts = [(w1 [I,H], w2 [H,I], w3 [I,H]), ...]."""
    b, t, H = x.shape
    xs = x.reshape(-1, H).astype(F32)
    idx, wts = route_top_k(ops.linear(xs, w_gate), top_k)
    if ROUTING_LOG is not None:
        ROUTING_LOG.append(idx.copy())                            # one [rows, top_k] entry per MoE block call, in layer order
    out = np.zeros_like(xs)
    for e, (w1, w2, w3) in enumerate(experts):
        rows, slot = np.nonzero(idx == e)
        if rows.size == 0:
            continue
        cur = xs[rows]
        y = ops.linear(ops.silu(ops.linear(cur, w1)) * ops.linear(cur, w3), w2)
        y = (y * wts[rows, slot][:, None]).astype(F32)
        np.add.at(out, rows, y)                                   # index_add
    return out.reshape(b, t, H)
