#!/bin/bash
# round-2 GPU session 10: dynamic tile scheduler + grouped colsum: parity, then epoch A/B static vs dynamic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== gemm + learner + agent parity"; timeout 900 python -m pytest tests/test_gpu_gemm_tc5s.py tests/test_gpu_learner.py tests/test_gpu_agent.py tests/test_gpu_mcp.py -q -p no:cacheprovider 2>&1 | tail -15
  for v in static dynamic; do
    echo "== bench sched=$v"
    PHC_TC5S_SCHED=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2> gpurun_out/s10_bench_$v.err | tee gpurun_out/s10_bench_$v.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'gemm',d['roofline_gemm']['achieved'],d['roofline_gemm']['forward_us'],d['roofline_gemm']['backward_us'])"
  done
  echo "== gemm microbench (dynamic)"; timeout 300 python tools/bench_gemm.py 20 s1
} > gpurun_out/s10.log 2>&1
cat gpurun_out/s10.log
