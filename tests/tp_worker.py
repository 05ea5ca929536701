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
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        json.dump({"ok": all(flags)}, open(out_path, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    (gpu_main if sys.argv[1] == "gpu" else cpu_main)(sys.argv[2])
