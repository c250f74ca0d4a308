#!/bin/bash
# round-2 GPU session 4: gemm_tc5s with warp-uniform issue code (elect.sync); env step with the provably-uniform warp index
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== tc5s parity"; timeout 600 python -m pytest tests/test_gpu_gemm_tc5s.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -6
  echo "== gemm microbench"; timeout 400 python tools/bench_gemm.py 20 persist,s1,s2
  echo "== env step parity"; timeout 600 python -m pytest tests/test_gpu_env_step.py tests/test_gpu_agent.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
  echo "== env kernel timing: new (uniform warp index)"; python tools/time_env.py 4096 60; python tools/time_env.py 16384 40
  echo "== env kernel timing: r1 form"; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_r1warp/libphc_b200.so python tools/time_env.py 4096 60; PHC_LIB_PATH=$PWD/phc_b200/lib/alt_r1warp/libphc_b200.so python tools/time_env.py 16384 40
  echo "== learner parity (tc5s)"; timeout 900 python -m pytest tests/test_gpu_learner.py -q -m gpu -p no:cacheprovider -k tc5s 2>&1 | tail -4
  echo "== bench tc5s pair"; PHC_GEMM=tc5s PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points 2>&1 | grep -E "phase_ms|value arm|Error|error|roofline_gemm"
  echo "== bench tc5s 1cta"; PHC_GEMM=tc5s PHC_TC5S_CTAS=1 PHC_PHASE_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-points 2>&1 | grep -E "phase_ms|value arm|Error|error|roofline_gemm"
} > gpurun_out/s4.log 2>&1
for v in s1 s2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 3 -c 3 -o gpurun_out/gemm_r2c_$v -f python tools/profile_gemm.py $v > gpurun_out/s4_ncu_$v.log 2>&1
done
ls -la gpurun_out | tail -4
cat gpurun_out/s4.log
