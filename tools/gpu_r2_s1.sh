#!/bin/bash
# round-2 GPU session 1: baseline evidence for the learner GEMMs (ncu counters), the >32-body kernels on hardware, phase breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== wide kernels on hardware"; PHC_TEST_WIDE=1 timeout 300 python -m pytest tests/test_gpu_wide.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15
  echo "== gemm microbench"; timeout 300 python tools/bench_gemm.py 20
  echo "== phase breakdown"; PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
} > gpurun_out/s1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -c 24 -o gpurun_out/gemm_r2_base_persist -f python tools/bench_gemm.py 1 persist > gpurun_out/s1_ncu_persist.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -c 24 -o gpurun_out/gemm_r2_base_plain -f python tools/bench_gemm.py 1 plain > gpurun_out/s1_ncu_plain.log 2>&1
ls -la gpurun_out | tail -8
tail -40 gpurun_out/s1.log
