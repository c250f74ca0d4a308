#!/bin/bash
# round-2 GPU session 25: template kernel with the first use of the scalar / dof loads deferred behind the tracking errors (pose-cache
# launches of the plain instantiations: H1, generic SMPL) -- GPU suite, A/B on the H1 workload and on the generic SMPL instantiation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
h1() {
  timeout 600 python bench.py --workload h1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('h1 value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['kernel_us'])"
}
{
  echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
  echo "== H1, deferred (default build)"; h1
  echo "== H1, not deferred"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_nodefer/libphc_b200.so h1
  echo "== generic SMPL instantiation (PHC_ENV_FAST=0), deferred"; PHC_ENV_FAST=0 python tools/time_env.py 4096 60
  echo "== generic SMPL instantiation (PHC_ENV_FAST=0), not deferred"; PHC_ENV_FAST=0 PHC_LIB_PATH=$PWD/phc_b200/lib/alt_nodefer/libphc_b200.so python tools/time_env.py 4096 60
  echo "== default (env_step_fast_kernel)"; python tools/time_env.py 4096 60
} > gpurun_out/s25.log 2>&1
cat gpurun_out/s25.log
