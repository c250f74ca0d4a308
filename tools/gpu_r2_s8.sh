#!/bin/bash
# round-2 GPU session 8: pre-split weight operand (PhcGemmDesc.B_lo) A/B -- parity, microbench, epoch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== tc5s parity"; timeout 600 python -m pytest tests/test_gpu_gemm_tc5s.py -q -p no:cacheprovider 2>&1 | tail -15
  echo "== gemm microbench"; timeout 300 python tools/bench_gemm.py 20 s1,s1p,s2p
  echo "== learner + agent parity"; timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_agent.py tests/test_gpu_mcp.py -q -p no:cacheprovider 2>&1 | tail -15
  for ps in 0 1; do
    echo "== bench presplit=$ps"
    PHC_TC5S_PRESPLIT=$ps timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-points --no-extras 2> gpurun_out/s8_bench_$ps.err | tee gpurun_out/s8_bench_$ps.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'gemm',d['roofline_gemm']['achieved'],d['roofline_gemm']['forward_us'],d['roofline_gemm']['backward_us'])"
  done
} > gpurun_out/s8.log 2>&1
cat gpurun_out/s8.log
