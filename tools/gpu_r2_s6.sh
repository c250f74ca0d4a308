#!/bin/bash
# round-2 GPU session 6: new parity tests, env-step launch-shape A/B, full default bench line (extras + points)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== new agent tests"; timeout 600 python -m pytest tests/test_gpu_agent.py -q -m gpu -p no:cacheprovider 2>&1 | tail -12
  echo "== env kernel timing: 4 warps/CTA (default)"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40
  for v in w7 w14; do
    echo "== env kernel timing: $v"
    PHC_LIB_PATH=$PWD/phc_b200/lib/alt_$v/libphc_b200.so timeout 200 python -m pytest tests/test_gpu_env_step.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
    PHC_LIB_PATH=$PWD/phc_b200/lib/alt_$v/libphc_b200.so python tools/time_env.py 4096 60; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_$v/libphc_b200.so python tools/time_env.py 16384 40
  done
  echo "== default bench"; SECONDS=0; timeout 900 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; echo "rc=$? wall=${SECONDS}s"; tail -12 gpurun_out/bench_r2_default.err; cat gpurun_out/bench_r2_default.json
} > gpurun_out/s6.log 2>&1
cat gpurun_out/s6.log
