#!/bin/bash
# round-2 GPU session 17: packed env-step kernel (4 envs per 3 warps) -- parity, A/B against the one-warp-per-env kernel, timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest env/agent/getup/robot/dropin (packed default)"; timeout 900 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_agent.py tests/test_gpu_getup.py tests/test_gpu_robot.py tests/test_gpu_dropin_construct.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -6
  echo "== time_env packed"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40; python tools/time_env.py 65536 20
  echo "== time_env PHC_ENV_PACKED=0"; PHC_ENV_PACKED=0 python tools/time_env.py 4096 60; PHC_ENV_PACKED=0 python tools/time_env.py 16384 40
  echo "== time_env PHC_ENV_PACKED=0 PHC_ENV_PDL=0"; PHC_ENV_PACKED=0 PHC_ENV_PDL=0 python tools/time_env.py 4096 60
  echo "== time_env packed PHC_ENV_PDL=0"; PHC_ENV_PDL=0 python tools/time_env.py 4096 60
  echo "== timeline 4096 packed"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_tl/libphc_b200.so python tools/timeline_env.py 4096 | head -34
} > gpurun_out/s17.log 2>&1
cat gpurun_out/s17.log
