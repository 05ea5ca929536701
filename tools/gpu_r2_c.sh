#!/bin/bash
# multi-GPU validation: N-rank tests + the default bench under torchrun (parity legs inside) -- run with gpurun --gpus N
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/c${N}_smi.txt 2>&1
timeout 1200 python -m pytest tests/test_tp.py -m gpu -q -rA --timeout 900 > $O/c${N}_pytest_tp.log 2>&1; echo "pytest rc=$?" >> $O/c${N}_pytest_tp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 \
   > $O/c${N}_bench.json 2> $O/c${N}_bench.err; echo "rc=$?" >> $O/c${N}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus $N --steps 10 --warmup 2 \
   > $O/c${N}_bench_ref.json 2> $O/c${N}_bench_ref.err; echo "rc=$?" >> $O/c${N}_bench_ref.err
