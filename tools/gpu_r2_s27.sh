#!/bin/bash
# round-2 GPU session 27: does the env-step microbenchmark inside bench.py see a depressed SM clock right after the power-capped epochs?
# same box, same process layout: settle 0 s vs 1 s before the roofline loop
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-points --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['kernel_us'],'clocks',d['clocks']['sm_mhz'])"
}
{
  echo "== settle 0"; PHC_ROOFLINE_SETTLE_S=0 run
  echo "== settle 1"; PHC_ROOFLINE_SETTLE_S=1 run
  echo "== settle 0"; PHC_ROOFLINE_SETTLE_S=0 run
  echo "== settle 1"; PHC_ROOFLINE_SETTLE_S=1 run
  echo "== standalone"; python tools/time_env.py 4096 60
} > gpurun_out/s27.log 2>&1
cat gpurun_out/s27.log
