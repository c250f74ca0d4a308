#!/bin/bash
# round-2 GPU session 13: vectorised normalise kernels: parity, epoch, phases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== learner + agent + mcp parity"; timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_agent.py tests/test_gpu_mcp.py tests/test_gpu_ppo_scalars.py -q -p no:cacheprovider 2>&1 | tail -12
  echo "== bench"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2> gpurun_out/s13_bench.err | tee gpurun_out/s13_bench.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'gemm',d['roofline_gemm']['achieved'],d['roofline_gemm']['forward_us'],d['roofline_gemm']['backward_us'],'launches',d['gpu_launches'])"
  echo "== phase breakdown (eager rollout, CUDA-event phases)"
  PHC_PHASE_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>&1 | grep -E "phase_ms|value arm"
} > gpurun_out/s13.log 2>&1
cat gpurun_out/s13.log
