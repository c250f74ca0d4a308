#!/bin/bash
# round-2 GPU session 12: deferred critic parity + A/B, dW split policy sweep with the wide tiles, phases, ncu counters of both tile shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run_bench() {   # label, env...
  local label=$1; shift
  echo "== bench $label"
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2> gpurun_out/s12_bench_$label.err | tee gpurun_out/s12_bench_$label.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'gemm',d['roofline_gemm']['achieved'],d['roofline_gemm']['forward_us'],d['roofline_gemm']['backward_us'])"
}
{
  echo "== agent parity"; timeout 900 python -m pytest tests/test_gpu_agent.py tests/test_gpu_learner.py -q -p no:cacheprovider 2>&1 | tail -12
  run_bench defer1 PHC_DEFER_CRITIC=1
  run_bench defer0 PHC_DEFER_CRITIC=0
  run_bench kb16 PHC_DW_KB_PER_SPLIT=16
  run_bench kb64 PHC_DW_KB_PER_SPLIT=64
  echo "== phase breakdown (eager rollout, CUDA-event phases)"
  PHC_PHASE_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>&1 | grep -E "phase_ms|value arm"
} > gpurun_out/s12.log 2>&1
ls -la gpurun_out | tail -4
cat gpurun_out/s12.log
