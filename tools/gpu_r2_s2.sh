#!/bin/bash
# round-2 GPU session 2: first run of the split-in-shared-memory GEMM (gemm_tc5s.cu): parity, microbench against the r1 kernel, ncu counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== tc5s parity"; timeout 600 python -m pytest tests/test_gpu_gemm_tc5s.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -25
  echo "== gemm microbench"; timeout 400 python tools/bench_gemm.py 20 persist,s1,s2
} > gpurun_out/s2.log 2>&1
for v in persist s1 s2; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 3 -c 3 -o gpurun_out/gemm_r2_$v -f python tools/profile_gemm.py $v > gpurun_out/s2_ncu_$v.log 2>&1
done
ls -la gpurun_out | tail -8
tail -45 gpurun_out/s2.log
