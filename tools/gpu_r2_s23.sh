#!/bin/bash
# round-2 GPU session 23: smoke() of the final build, CUDA-event phase breakdown of one epoch of the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
  echo "== phase breakdown (eager rollout, CUDA-event phases)"
  PHC_PHASE_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>&1 | grep -E "phase_ms|value arm|^\{" | cut -c1-1500
  echo "== same box, graph rollout"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'],'gemm',d['roofline_gemm']['achieved'])"
} > gpurun_out/s23.log 2>&1
cat gpurun_out/s23.log
