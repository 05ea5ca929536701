#!/bin/bash
# Dev tool: the multi-GPU bench lines of the BASELINE.json configs on N GPUs of one box (run under gpurun --gpus N).
N=${1:-2}
OUT=gpurun_out
run() { # name, args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N "$@" \
      > $OUT/bench_${name}_n$N.json 2> $OUT/bench_${name}_n$N.err
  echo "$name rc=$?"; grep -h '^{' $OUT/bench_${name}_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d.get('roofline') or {}
    print('  ', d['metric'], round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(r.get('frac') or 0,3), d.get('decode_after_prefill'))
"
}
run mistral_b1 --steps 128 --warmup 8
run qwen_prefill --workload qwen25_7b_prefill4k --steps 4
run qwen_b1 --workload qwen25_7b_b1 --steps 128
run mixtral_b32 --workload mixtral8x7b_b32 --steps 32 --warmup 4
[ "$N" -lt 8 ] && run minilm --workload minilm_256x128 --steps 20
