#!/bin/bash
# round-2 GPU session 24 (gpurun --gpus 2): the final build under torchrun -- 2-rank bench line (NCCL all-reduce on the side stream,
# PDL-launched env step inside the per-rank rollout graph)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-points --no-cpu-baseline 2> gpurun_out/s24_bench2.err | tee gpurun_out/bench_r2_final_2gpu.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'n',d['n_gpus'],'roofline',d['roofline']['frac'])"
  tail -3 gpurun_out/s24_bench2.err
} > gpurun_out/s24.log 2>&1
cat gpurun_out/s24.log
