#!/bin/bash
# round-2 GPU session 14 (run with gpurun --gpus 2): NCCL data-parallel check + 2-GPU bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== check_ddp (2 ranks)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_ddp.py 2>&1 | tail -8
  echo "== bench --gpus 2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras --no-points --no-cpu-baseline 2> gpurun_out/s14_bench2.err | tee gpurun_out/bench_r2_2gpu.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'n',d['n_gpus'])"
  tail -3 gpurun_out/s14_bench2.err
  echo "== bench --gpus 1 (same box)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-points --no-cpu-baseline 2> gpurun_out/s14_bench1.err | tee gpurun_out/s14_bench1.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'])"
} > gpurun_out/s14.log 2>&1
cat gpurun_out/s14.log
