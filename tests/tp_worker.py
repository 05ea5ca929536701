"""Worker for the tensor-parallel tests: launched once per rank by torch.distributed.run (backend nccl on GPUs, or the
numpy simulation of the same sharding over gloo on CPUs)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gpu_main(out_path):
    import torch
    import torch.distributed as dist
    from dataclasses import replace
    from fastllm_b200 import models, tp
    from helpers import golden_weights
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tp.init_tensor_parallel(rank, world, local)
    res = {}
    for name in ("mistral",):
        cfg, w, g = golden_weights(name)
        cf = models.ConfigFile(cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                               cfg.num_key_value_heads, cfg.rms_norm_eps, cfg.rope_theta, cfg.max_position_embeddings, cfg.sliding_window,
                               tp_rank=rank, tp_size=world)
        # every rank hands over the FULL tensors, exactly like load_model would; the library keeps its shard
        model, cache = models.MistralWithConfig.initialize_model(cf, w, "bf16", local)
        ids, logits = models.Model(model, cache, eos_token_id=None).generate(g["prompt"], len(g["faithful_ids"]), return_logits=True)
        # device-side synthetic init must agree between shard layouts as well
        model2, cache2 = models.MistralWithConfig.initialize_model(cf, None, "bf16", local, random_seed=int(g["seed"]), std=float(g["std"]))
        l2 = model2.forward(np.asarray(g["prompt"], dtype=np.uint32)[None], 0, cache2)
        # batch 3 goes through the dense (tcgen05) path with its own collectives
        c3 = models.DeviceCache(model.dev, 3, 64)
        p3 = np.stack([g["prompt"], g["prompt"][::-1], np.roll(g["prompt"], 3)]).astype(np.uint32)
        l3 = c3.forward(p3, 0)
        res[name] = {"ids": [int(i) for i in ids], "logits": np.stack(logits).tolist(), "synth_logits": l2[0, 0].tolist(), "batch3": l3.tolist()}
    # tiny Qwen2 (7 q heads, 1 kv head): at TP-2 the kv head is replicated and the query group padded to 4 + 4 heads
    cfg, w, g = golden_weights("qwen2")
    cf = models.ConfigFile(cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                           cfg.num_key_value_heads, cfg.rms_norm_eps, cfg.rope_theta, cfg.max_position_embeddings, cfg.sliding_window,
                           tp_rank=rank, tp_size=world)
    model, cache = models.QwenWithConfig.initialize_model(cf, w, "bf16", local)
    ids, logits = models.Model(model, cache, eos_token_id=None).generate(g["prompt"], len(g["faithful_ids"]), return_logits=True)
    model2, cache2 = models.QwenWithConfig.initialize_model(cf, None, "bf16", local, random_seed=int(g["seed"]), std=float(g["std"]))
    l2 = model2.forward(np.asarray(g["prompt"], dtype=np.uint32)[None], 0, cache2)
    c3 = models.DeviceCache(model.dev, 3, 64)
    p3 = np.stack([g["prompt"], g["prompt"][::-1], np.roll(g["prompt"], 3)]).astype(np.uint32)
    res["qwen2"] = {"ids": [int(i) for i in ids], "logits": np.stack(logits).tolist(), "synth_logits": l2[0, 0].tolist(),
                    "batch3": c3.forward(p3, 0).tolist(), "golden_ids": [int(i) for i in g["faithful_ids"]]}
    del c3, model, model2
    # true-width 2-layer Mistral-7B: batch-1 decode runs in the persistent kernel, whose all-reduce goes over NVLink peer memory
    cfw = models.ConfigFile(4096, 14336, 32000, 2, 32, 8, 1e-5, 10000.0, 256, 4096, tp_rank=rank, tp_size=world)
    mw, _ = models.MistralWithConfig.initialize_model(cfw, None, "bf16", local, random_seed=0, std=0.02)
    cw = models.DeviceCache(mw.dev, 1, 200)
    from oracle import synth
    pw = synth.token_ids(1, 32000, (1, 70))
    tok = cw.forward_greedy(pw, 0)
    step_ids, step_logits = [], []
    for s_ in range(5):
        lg = cw.forward(tok.reshape(1, 1), 70 + s_)                  # one persistent launch per step
        step_logits.append(lg[0].tolist())
        tok = np.array([models.sample_argmax(lg[0])], dtype=np.uint32)
        step_ids.append(int(tok[0]))
    loop_ids, _ = cw.decode_greedy_loop(tok, 75, 6)                   # six steps inside one launch
    res["wide"] = {"ids": step_ids, "logits": step_logits, "loop_ids": [int(i) for i in loop_ids[:, 0]]}
    del cw, mw
    # Mixtral: expert parallelism (experts sharded across ranks, attention replicated, one all-reduce per MoE block)
    cfg, w, g = golden_weights("mixtral")
    cf = models.ConfigFile(cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                           cfg.num_key_value_heads, cfg.rms_norm_eps, cfg.rope_theta, cfg.max_position_embeddings, cfg.sliding_window,
                           tp_rank=rank, tp_size=world, num_local_experts=cfg.num_local_experts, num_experts_per_tok=cfg.num_experts_per_tok)
    model, cache = models.MixtralWithConfig.initialize_model(cf, w, "bf16", local)
    ids, logits = models.Model(model, cache, eos_token_id=None).generate(g["prompt"], len(g["faithful_ids"]), return_logits=True)
    res["mixtral"] = {"ids": [int(i) for i in ids], "logits": np.stack(logits).tolist(), "golden_ids": [int(i) for i in g["faithful_ids"]]}
    # Mixtral, expert parallelism with data-parallel attention: every rank runs ITS slice of a 4-sequence batch, tokens travel to
    # the experts and back by all-to-all (dispatch / combine); prefill + two decode steps
    cf_dp = replace(cf, ep_dp_attention=True)
    mdp, _ = models.MixtralWithConfig.initialize_model(cf_dp, w, "bf16", local)
    p4 = np.stack([np.roll(g["prompt"], k) for k in range(4)]).astype(np.uint32)
    per = 4 // world
    mine = p4[rank * per:(rank + 1) * per]
    cdp = models.DeviceCache(mdp.dev, per, 64)
    outs = [cdp.forward(mine, 0)]
    for s_ in range(2):
        nxt = np.array([[models.sample_argmax(r)] for r in outs[-1]], dtype=np.uint32)
        outs.append(cdp.forward(nxt, mine.shape[1] + s_))
    gathered = [None] * world
    dist.all_gather_object(gathered, np.stack(outs).tolist())
    res["mixtral_dp"] = np.concatenate([np.array(x) for x in gathered], axis=1).tolist()      # [3 calls, 4 sequences, V]
    if rank == 0:
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


def gpu_wide_main(out_path):
    """True-width shapes on 1 / 2 / 4 / 8 ranks (tests/test_tp.py compares N ranks with 1 rank): 2-layer Mistral-7B (persistent kernel
    with the in-kernel NVLink all-reduce at batch 1, the NCCL dense path at batch 8 and in the prefill), 2-layer Qwen2.5-7B with its
    REAL 28 q / 4 kv head layout and 152064-row vocabulary (at 8 ranks: every kv head on two ranks, query groups dealt 4 + 3 + one
    zero head), and a 1-layer Mixtral-8x7B (8 experts, top-2) expert-parallel with data-parallel attention."""
    import torch
    import torch.distributed as dist
    from fastllm_b200 import models, tp
    from oracle import synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tp.init_tensor_parallel(rank, world, local)
    res = {}

    def dense_lm(cls, cf, vocab, tag):
        # every multi-step comparison is TEACHER-FORCED with fixed synthetic tokens: random-init logits have top-2 gaps down to
        # 1e-3, and one flipped arg-max would turn a summation-order difference into a different continuation
        m, _ = cls.initialize_model(cf, None, "bf16", local, random_seed=0, std=0.02)
        c = models.DeviceCache(m.dev, 1, 200)
        p = synth.token_ids(1, vocab, (1, 70))
        feed = synth.token_ids(4, vocab, (5,))
        rows = [c.forward(p, 0)[0]]                                       # 70-token prefill: dense path, NCCL all-reduces
        for s_ in range(5):                                               # one persistent launch per step
            rows.append(c.forward(feed[s_].reshape(1, 1), 70 + s_)[0])
        first = np.array([feed[4]], dtype=np.uint32)
        loop_ids, _ = c.decode_greedy_loop(first, 75, 6)                  # six steps inside one launch, greedy feedback on device
        c2 = models.DeviceCache(m.dev, 1, 200)                            # the loop's ids must be the arg-max of step-by-step logits
        c2.forward(p, 0)
        for s_ in range(5):
            c2.forward(feed[s_].reshape(1, 1), 70 + s_)
        tok, slack = int(feed[4]), []
        for s_ in range(6):
            lg = c2.forward(np.array([[tok]], dtype=np.uint32), 75 + s_)[0]
            tok = int(loop_ids[s_, 0])
            slack.append(float(lg.max() - lg[tok]))
        c8 = models.DeviceCache(m.dev, 8, 128)                            # batch 8: tcgen05 dense decode, NCCL collectives in the graph
        p8 = synth.token_ids(2, vocab, (8, 20))
        f8 = synth.token_ids(5, vocab, (3, 8, 1))
        l8 = [c8.forward(p8, 0)]
        for s_ in range(3):
            l8.append(c8.forward(f8[s_], 20 + s_))

        def top(r):      # (arg-max, top-1/top-2 gap) of a logits row
            t2 = np.partition(r, -2)[-2:]
            return [int(models.sample_argmax(r)), float(t2[1] - t2[0])]

        res[tag] = {"top": [top(r) for r in rows], "logits": [r[:4096].tolist() for r in rows], "loop_slack": slack,
                    "b8_top": [[top(r) for r in l] for l in l8], "b8_logits": [l[:, :2048].tolist() for l in l8]}

    dense_lm(models.MistralWithConfig, models.ConfigFile(4096, 14336, 32000, 2, 32, 8, 1e-5, 10000.0, 256, 4096, tp_rank=rank, tp_size=world),
             32000, "mistral7b")
    dense_lm(models.QwenWithConfig, models.ConfigFile(3584, 18944, 152064, 2, 28, 4, 1e-6, 1000000.0, 256, None, tp_rank=rank, tp_size=world),
             152064, "qwen25_7b")
    # Mixtral-8x7B shapes, one layer: 8 sequences dealt out to the ranks, tokens travel to the experts and back by all-to-all
    cf = models.ConfigFile(4096, 14336, 32000, 1, 32, 8, 1e-5, 1000000.0, 256, 4096, tp_rank=rank, tp_size=world, num_local_experts=8,
                           num_experts_per_tok=2, ep_dp_attention=world > 1)
    m, _ = models.MixtralWithConfig.initialize_model(cf, None, "bf16", local, random_seed=0, std=0.02)
    p8 = synth.token_ids(3, 32000, (8, 12))
    per = 8 // world
    mine = p8[rank * per:(rank + 1) * per]
    c = models.DeviceCache(m.dev, per, 64)
    fm = synth.token_ids(6, 32000, (3, 8, 1))
    outs = [c.forward(mine, 0)]
    for s_ in range(3):
        outs.append(c.forward(fm[s_][rank * per:(rank + 1) * per], 12 + s_))
    gathered = [None] * world
    dist.all_gather_object(gathered, np.stack(outs)[:, :, :2048].tolist())
    allo = np.concatenate([np.array(x) for x in gathered], axis=1)                       # [4 calls, 8 sequences, 2048]
    res["mixtral"] = {"logits": allo.tolist()}
    if rank == 0:
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


def cpu_main(out_path):
    """No GPU: simulate the library's TP data flow in numpy with the SAME shard windows, all-reduce over gloo."""
    import torch
    import torch.distributed as dist
    from fastllm_b200 import tp
    from helpers import golden_weights
    from oracle import candle_ops as ops
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    cfg, w, g = golden_weights("mistral")
    ws = tp.shard_weights(w, cfg.num_attention_heads, cfg.num_key_value_heads, rank, world)
    x = np.random.default_rng(0).standard_normal((5, cfg.hidden_size)).astype(np.float32)
    p = "model.layers.0."
    # row-parallel pair: (column-parallel gate/up -> SiLU*up) feeds the local slice of down_proj; partial sums all-reduced
    act = ops.silu(ops.linear(x, ws[p + "mlp.gate_proj.weight"])) * ops.linear(x, ws[p + "mlp.up_proj.weight"])
    part = torch.from_numpy(ops.linear(act, ws[p + "mlp.down_proj.weight"]))
    dist.all_reduce(part)
    full = ops.linear(ops.silu(ops.linear(x, w[p + "mlp.gate_proj.weight"])) * ops.linear(x, w[p + "mlp.up_proj.weight"]),
                      w[p + "mlp.down_proj.weight"])
    # vocab-parallel head: all-gather of the local logits slices, rank-major
    loc = torch.from_numpy(ops.linear(x, ws["lm_head.weight"]))
    parts = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc)
    logits = np.concatenate([t.numpy() for t in parts], axis=1)
    ok = bool(np.abs(part.numpy() - full).max() < 1e-4 and np.abs(logits - ops.linear(x, w["lm_head.weight"])).max() < 1e-4)
    # q/k/v are split by head: the local q rows are whole heads
    d = cfg.head_dim
    ok = ok and ws[p + "self_attn.q_proj.weight"].shape[0] == cfg.num_attention_heads // world * d
    ok = ok and ws[p + "self_attn.k_proj.weight"].shape[0] == cfg.num_key_value_heads // world * d
    # more ranks than kv heads (tiny Qwen2: 7 q heads, 1 kv head at TP-2 -- the Qwen2.5-7B TP-8 situation): the kv head is
    # replicated, the query group is dealt 4 + 3 heads; attention per local head + row-parallel o_proj, all-reduced, must equal
    # the unsharded attention block
    qcfg, qw, _ = golden_weights("qwen2")
    nh, nkv, dq = qcfg.num_attention_heads, qcfg.num_key_value_heads, qcfg.head_dim
    qs = tp.shard_weights(qw, nh, nkv, rank, world)
    q0, qn, nhl, k0, kn = tp.head_layout(nh, nkv, rank, world)
    ok = ok and (q0, qn, nhl, k0, kn) == ((0, 4, 4, 0, 1) if rank == 0 else (4, 3, 4, 0, 1))
    xq = np.random.default_rng(1).standard_normal((6, qcfg.hidden_size)).astype(np.float32)

    def attn_block(wts, heads, kvheads):
        q = (ops.linear(xq, wts[p + "self_attn.q_proj.weight"]) + wts[p + "self_attn.q_proj.bias"]).reshape(6, heads, dq)
        k = (ops.linear(xq, wts[p + "self_attn.k_proj.weight"]) + wts[p + "self_attn.k_proj.bias"]).reshape(6, kvheads, dq)
        v = (ops.linear(xq, wts[p + "self_attn.v_proj.weight"]) + wts[p + "self_attn.v_proj.bias"]).reshape(6, kvheads, dq)
        outs = []
        for h in range(heads):
            kv = 0          # one kv head in this config (replicated on both ranks)
            sc = q[:, h] @ k[:, kv].T / np.sqrt(dq)
            sc = np.where(np.tril(np.ones((6, 6), dtype=bool)), sc, -np.inf)
            outs.append(ops.softmax_last_dim(sc.astype(np.float32)) @ v[:, kv])
        return ops.linear(np.concatenate(outs, axis=1).astype(np.float32), wts[p + "self_attn.o_proj.weight"])

    part_a = torch.from_numpy(attn_block(qs, qn, kn))
    dist.all_reduce(part_a)
    ok = ok and bool(np.abs(part_a.numpy() - attn_block(qw, nh, nkv)).max() < 1e-4)
    # Mixtral expert parallelism with data-parallel attention (fl_config.ep_dp_attention): every rank owns a slice of the rows and
    # E / world experts.  DISPATCH: the rows (+ their routing weights) travel to every expert rank; every rank runs its experts over
    # ALL rows with the masked routing weight (zero when the expert was not selected); COMBINE: the partial outputs of rank p's rows
    # travel back to rank p, which sums them in rank order.  Must equal the oracle's sparse-MoE block on the owner's rows.
    from oracle import mixtral as omix
    mcfg, mw, _ = golden_weights("mixtral")
    E, topk = mcfg.num_local_experts, mcfg.num_experts_per_tok
    pm = "model.layers.0.block_sparse_moe."
    xm = np.random.default_rng(2).standard_normal((4 * world, mcfg.hidden_size)).astype(np.float32)
    per = xm.shape[0] // world
    mine = xm[rank * per:(rank + 1) * per]
    idx, wts = omix.route_top_k(ops.linear(mine, mw[pm + "gate.weight"]), topk)
    route = np.zeros((per, E), dtype=np.float32)
    for r_ in range(per):
        route[r_, idx[r_]] = wts[r_]
    g_rows = [torch.empty(per, mcfg.hidden_size) for _ in range(world)]
    g_route = [torch.empty(per, E) for _ in range(world)]
    dist.all_gather(g_rows, torch.from_numpy(np.ascontiguousarray(mine)))          # dispatch (every rank receives every block)
    dist.all_gather(g_route, torch.from_numpy(route))
    rows_all = np.concatenate([t.numpy() for t in g_rows]); route_all = np.concatenate([t.numpy() for t in g_route])
    e_local = E // world
    part = np.zeros_like(rows_all)
    for e in range(rank * e_local, (rank + 1) * e_local):
        w1, w2, w3 = (mw[pm + f"experts.{e}.{n}.weight"] for n in ("w1", "w2", "w3"))
        y = ops.linear(ops.silu(ops.linear(rows_all, w1)) * ops.linear(rows_all, w3), w2)
        part += route_all[:, e:e + 1] * y
    back = [torch.empty(world * per, mcfg.hidden_size) for _ in range(world)]
    dist.all_gather(back, torch.from_numpy(part))                                   # combine: keep block `rank` of every source
    combined = np.zeros((per, mcfg.hidden_size), dtype=np.float32)
    for src in range(world):                                                        # fixed rank order
        combined += back[src].numpy()[rank * per:(rank + 1) * per]
    experts = [(mw[pm + f"experts.{e}.w1.weight"], mw[pm + f"experts.{e}.w2.weight"], mw[pm + f"experts.{e}.w3.weight"]) for e in range(E)]
    want = omix.sparse_moe_block(mine[None], mw[pm + "gate.weight"], experts, topk)[0]
    ok = ok and bool(np.abs(combined - want).max() < 1e-4)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        json.dump({"ok": all(flags)}, open(out_path, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    {"gpu": gpu_main, "gpu_wide": gpu_wide_main, "cpu": cpu_main}[sys.argv[1]](sys.argv[2])
