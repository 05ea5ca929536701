"""Dev tool: phase times of the persistent kernel under tensor parallelism (run with torch.distributed.run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from dataclasses import replace
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from fastllm_b200 import models, presets, tp
tp.init_tensor_parallel(rank, world, local)
cls, cf = presets.PRESETS["mistral7b"]
cf = replace(cf, tp_rank=rank, tp_size=world)
model, _ = cls.initialize_model(cf, None, "bf16", local, random_seed=0)
cache = models.DeviceCache(model.dev, 1, 2200)
cache.fill_synthetic(1, 2048)
if rank == 0:
    os.environ["FL_PK_DEBUG"] = "1"
for i in range(3):
    cache.forward(np.array([[5]], dtype=np.uint32), 2048 + i)
dist.barrier()
