#!/bin/bash
# round-2 GPU session 5: full GPU suite on the tc5s default + device ring head + graph-replayed rollout + pipelined minibatches; bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15
  echo "== bench (graph rollout)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-points 2>&1 | grep -vE "^\[bench" 
  echo "== bench eager rollout + phases"; PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points 2>&1 | grep -E "phase_ms|value arm|Error|error"
} > gpurun_out/s5.log 2>&1
cat gpurun_out/s5.log
