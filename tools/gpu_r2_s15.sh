#!/bin/bash
# round-2 GPU session 15: full GPU test suite of HEAD, env-step timeline + launch / load floors, env timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
  echo "== time_env default"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40
  echo "== timeline 4096"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_tl/libphc_b200.so python tools/timeline_env.py 4096
  echo "== timeline 16384"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_tl/libphc_b200.so python tools/timeline_env.py 16384 | head -30
  echo "== floor: empty kernel (exit1)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_exit1/libphc_b200.so python tools/time_env.py 4096 60
  echo "== floor: loads only (exit2)"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_exit2/libphc_b200.so python tools/time_env.py 4096 60
  PHC_LIB_PATH=$PWD/phc_b200/lib/alt_exit2/libphc_b200.so python tools/time_env.py 16384 40
} > gpurun_out/s15.log 2>&1
cat gpurun_out/s15.log
