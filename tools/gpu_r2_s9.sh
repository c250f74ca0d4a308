#!/bin/bash
# round-2 GPU session 9: stage-layout A/B of the GEMM on one box, fused-rms traceback, phase breakdown of the current epoch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== failing tests, full traceback"; timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_agent.py tests/test_gpu_dropin_construct.py -q -p no:cacheprovider -k "fused_rms or reset_then_step or parse_task or eval_sweep" --tb=long 2>&1 | tail -60
  echo "== tc5s parity on the ring-layout build"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_blo_ring/libphc_b200.so timeout 600 python -m pytest tests/test_gpu_gemm_tc5s.py -q -p no:cacheprovider 2>&1 | tail -4
  echo "== gemm microbench, default build (B lo in stage, 3 stages)"; timeout 300 python tools/bench_gemm.py 20 s1,s1p
  echo "== gemm microbench, ring-layout build (4 raw stages)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_blo_ring/libphc_b200.so timeout 300 python tools/bench_gemm.py 20 s1
  for v in default_presplit default_nopresplit ring; do
    echo "== bench $v"
    case $v in
      default_presplit) export PHC_TC5S_PRESPLIT=1; unset PHC_LIB_PATH;;
      default_nopresplit) export PHC_TC5S_PRESPLIT=0; unset PHC_LIB_PATH;;
      ring) export PHC_TC5S_PRESPLIT=0; export PHC_LIB_PATH=$PWD/phc_b200/lib/alt_blo_ring/libphc_b200.so;;
    esac
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2> gpurun_out/s9_bench_$v.err | tee gpurun_out/s9_bench_$v.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'gemm',d['roofline_gemm']['achieved'],d['roofline_gemm']['forward_us'],d['roofline_gemm']['backward_us'])"
  done
  unset PHC_LIB_PATH PHC_TC5S_PRESPLIT
  echo "== phase breakdown (eager rollout, CUDA-event phases)"
  PHC_PHASE_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>&1 | grep -E "phase_ms|value arm"
} > gpurun_out/s9.log 2>&1
cat gpurun_out/s9.log
