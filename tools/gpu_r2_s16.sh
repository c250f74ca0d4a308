#!/bin/bash
# round-2 GPU session 16: PDL launch of the env step (griddepcontrol.wait + phc_l2_flush), scalars-first request order
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest env/agent (default build, PDL on)"; timeout 600 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_agent.py tests/test_gpu_getup.py tests/test_gpu_robot.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
  echo "== time_env default (PDL)"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40
  echo "== time_env PHC_ENV_PDL=0"; PHC_ENV_PDL=0 python tools/time_env.py 4096 60
  echo "== scalars first"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_sf/libphc_b200.so python tools/time_env.py 4096 60; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_sf/libphc_b200.so python tools/time_env.py 16384 40
  echo "== timeline 4096 (PDL)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_tl/libphc_b200.so python tools/timeline_env.py 4096 | head -34
  echo "== timeline 4096 (PDL, scalars first)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_sftl/libphc_b200.so python tools/timeline_env.py 4096 | head -34
  echo "== floor: empty kernel (exit1)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_exit1/libphc_b200.so python tools/time_env.py 4096 60
  echo "== floor: loads only (exit2)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_exit2/libphc_b200.so python tools/time_env.py 4096 60
} > gpurun_out/s16.log 2>&1
cat gpurun_out/s16.log
