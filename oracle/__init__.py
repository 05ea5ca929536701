"""CPU oracle for the fastllm transformer forward pass.  TEST INFRASTRUCTURE ONLY.

This package is a CPU *restatement* (numpy, f32) of the arithmetic the
reference issues through candle 0.8.x for the hot path named in
BASELINE.json / SURVEY.md section 8:

  * Llama / Mistral / Qwen2 causal-LM forward (prefill + decode)
      reference adapters: src/models/llama.rs:147-149, src/models/mistral.rs:206-236,
      src/models/qwen.rs:123-151  ->  candle-transformers 0.8.x models::{llama,mistral,qwen2}
  * the hand-written BERT/MiniLM encoder        src/models/embeddings.rs:130-447
  * the greedy generate loop + LogitsProcessor  src/models/mod.rs:363-463
  * the Mixtral sparse-MoE block (candle-transformers models::mixtral; not wired in the reference)

PARITY STATUS: **parity unpinned by the reference**.  The reference holds no
golden vectors or known-answer tests for this path, its
arithmetic lives in the un-vendored crates candle-core/candle-nn/
candle-transformers "^0.8.2" (Cargo.toml:19-21, no Cargo.lock), and there is no
Rust toolchain in the image, so the reference itself cannot be run here.  The
oracle is instead cross-checked against an *independent* implementation
(HuggingFace transformers 5.5, torch CPU f32) by tests/golden/make_golden.py,
and the resulting vectors are committed under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product path (fastllm_b200/) never
does; it fails loudly when the CUDA library is missing.
"""
