#!/bin/bash
# multi-GPU validation: N-rank tests + the default bench under torchrun (parity legs inside) -- run with gpurun --gpus N
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/c${N}_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_tp.py -m gpu -q -rA --timeout 900 -k "${2:-tp}" > $O/c${N}_pytest_tp.log 2>&1; echo "pytest rc=$?" >> $O/c${N}_pytest_tp.log
FL_BENCH_ENV_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 20 --warmup 5 \
   > $O/c${N}_bench.json 2> $O/c${N}_bench.err; echo "rc=$?" >> $O/c${N}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus $N --steps 10 --warmup 2 \
   > $O/c${N}_bench_ref.json 2> $O/c${N}_bench_ref.err; echo "rc=$?" >> $O/c${N}_bench_ref.err
FL_PK_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus $N --steps 8 --warmup 3 --no-extras --no-cpu \
   > $O/c${N}_pkdebug.json 2> $O/c${N}_pkdebug.err
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29580 bench.py --gpus 4 --steps 20 --warmup 5 \
     > $O/c4_bench.json 2> $O/c4_bench.err; echo "rc=$?" >> $O/c4_bench.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --workload qwen25_7b_prefill4k --steps 8 --warmup 3 \
     > $O/c8_qwen_prefill.json 2> $O/c8_qwen_prefill.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --workload minilm_256x128 --steps 20 --warmup 5 \
     > $O/c8_minilm.json 2> $O/c8_minilm.err
fi
