#!/bin/bash
# A/B of experiment builds of the fused env-step kernel on a GPU box (build them first, here, with
# `python -c "from phc_b200 import build as b; b.build_variant('w2', ['-DPHC_EXP_WARPS=2'])"`): parity first, then the
# L2-flushed CUDA-event timing of tools/time_env.py.  Results go to gpurun_out/ab_env.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== default (FAST specialisation)"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40
  echo "== default, PHC_ENV_FAST=0 (generic instantiation)"; PHC_ENV_FAST=0 python tools/time_env.py 4096 60
  for d in phc_b200/lib/alt_*/; do
    v=$(basename "$d")
    echo "== $v"
    PHC_LIB_PATH="$PWD/$d/libphc_b200.so" timeout 120 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_agent.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
    PHC_LIB_PATH="$PWD/$d/libphc_b200.so" python tools/time_env.py 4096 60
  done
} > gpurun_out/ab_env.log 2>&1
