#!/bin/bash
# round-2 GPU session 20: instruction diets of env_step_fast.cu, each with its parity tests: approximate division / square root
# (-prec-div=false -prec-sqrt=false), heading quaternion from half-angle identities, scalar requests ahead of the simulator block
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== default"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40
  for v in pd ha pdha sf pdhasf; do
    echo "== $v"
    PHC_LIB_PATH=$PWD/phc_b200/lib/alt_$v/libphc_b200.so timeout 300 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_agent.py -m gpu -q -p no:cacheprovider -k "oracle or golden or shipped or reset_then_step or specialised" 2>&1 | tail -3
    PHC_LIB_PATH=$PWD/phc_b200/lib/alt_$v/libphc_b200.so python tools/time_env.py 4096 60
    PHC_LIB_PATH=$PWD/phc_b200/lib/alt_$v/libphc_b200.so python tools/time_env.py 16384 40
  done
} > gpurun_out/s20.log 2>&1
cat gpurun_out/s20.log
