#!/bin/bash
# round-2 GPU session 3: gemm_tc5s with the decoupled raw / lo rings + ReLU bit masks; grouped learner path end to end
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== tc5s parity"; timeout 600 python -m pytest tests/test_gpu_gemm_tc5s.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15
  echo "== learner parity (all back ends incl. grouped tc5s)"; timeout 900 python -m pytest tests/test_gpu_learner.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15
  echo "== gemm microbench"; timeout 400 python tools/bench_gemm.py 20 persist,s1,s2
  echo "== bench tc5 (r1 path)"; PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "phase_ms|value arm"
  echo "== bench tc5s pair"; PHC_GEMM=tc5s PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "phase_ms|value arm|Error|error"
  echo "== bench tc5s 1cta"; PHC_GEMM=tc5s PHC_TC5S_CTAS=1 PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "phase_ms|value arm|Error|error"
  echo "== agent tests with tc5s"; PHC_GEMM=tc5s timeout 600 python -m pytest tests/test_gpu_agent.py tests/test_gpu_mcp.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8
} > gpurun_out/s3.log 2>&1
for v in s1 s2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 3 -c 3 -o gpurun_out/gemm_r2b_$v -f python tools/profile_gemm.py $v > gpurun_out/s3_ncu_$v.log 2>&1
done
ls -la gpurun_out | tail -6
cat gpurun_out/s3.log
