#!/bin/bash
# round-2 GPU session 19: env_step_fast.cu with the first use of the scalar / dof loads moved behind the self observation -- parity, A/B, timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest env/agent/getup (fast kernel default)"; timeout 900 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_agent.py tests/test_gpu_getup.py tests/test_gpu_dropin_construct.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
  echo "== time_env fast kernel"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40; python tools/time_env.py 65536 20
  echo "== time_env PHC_ENV_FASTK=0"; PHC_ENV_FASTK=0 python tools/time_env.py 4096 60; PHC_ENV_FASTK=0 python tools/time_env.py 16384 40
  echo "== time_env cache late"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_cl/libphc_b200.so python tools/time_env.py 4096 60; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_cl/libphc_b200.so python tools/time_env.py 16384 40
  echo "== timeline 4096 fast kernel"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_tl/libphc_b200.so python tools/timeline_env.py 4096 | head -34
  echo "== timeline 4096 fast kernel, cache late"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_cltl/libphc_b200.so python tools/timeline_env.py 4096 | head -34
} > gpurun_out/s19.log 2>&1
cat gpurun_out/s19.log
