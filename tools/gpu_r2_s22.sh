#!/bin/bash
# round-2 GPU session 22: the whole GPU test suite and smoke() on the final build, then the ncu launch list of agent construction +
# the rollout + the first minibatches of one epoch (eager rollout so that every kernel is a launch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
  echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
  echo "== ncu launch list"
  PHC_GRAPH_ROLLOUT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4200 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-points --no-extras > gpurun_out/s22_ncu_bench.log 2>&1
  tail -2 gpurun_out/s22_ncu_bench.log; wc -l gpurun_out/launches_r2b.csv; gzip -f gpurun_out/launches_r2b.csv
} > gpurun_out/s22.log 2>&1
cat gpurun_out/s22.log
