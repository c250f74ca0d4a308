#!/bin/bash
# round-2 GPU session 11: wide-tile GEMM (128 x 256 x 16, gemm_tc5w.cu): parity, microbench, epoch A/B against the 128 x 128 x 32 kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== gemm parity (all three tile configurations)"; timeout 900 python -m pytest tests/test_gpu_gemm_tc5s.py -q -p no:cacheprovider -x 2>&1 | tail -25
  echo "== gemm microbench"; timeout 300 python tools/bench_gemm.py 20 s1,w
  echo "== learner + agent parity (wide default)"; timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_agent.py tests/test_gpu_mcp.py -q -p no:cacheprovider 2>&1 | tail -8
  for v in 128 256; do
    echo "== bench tile=$v"
    PHC_TC5_TILE=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2> gpurun_out/s11_bench_$v.err | tee gpurun_out/s11_bench_$v.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'gemm',d['roofline_gemm']['achieved'],d['roofline_gemm']['forward_us'],d['roofline_gemm']['backward_us'])"
  done
} > gpurun_out/s11.log 2>&1
cat gpurun_out/s11.log
