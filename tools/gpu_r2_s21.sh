#!/bin/bash
# round-2 GPU session 21: ncu --set full capture of env_step_fast_kernel (-> traffic json), the whole GPU test suite, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== ncu --set full env_step_fast_kernel"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_fast_kernel -s 6 -c 1 -f -o gpurun_out/env_step_r2 python tools/profile_env.py 4096 12 2>&1 | tail -4
  python tools/env_traffic_from_ncu.py gpurun_out/env_step_r2.ncu-rep 4096 profiles/env_step_traffic.json gpurun_out/env_step_traffic.json
  echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
  echo "== bench.py (default)"; timeout 1200 python bench.py 2> gpurun_out/s21_bench.err | tee gpurun_out/bench_r2_final_1gpu.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['kernel_us'],[p['frac'] for p in d['roofline'].get('points',[])],'gemm',d['roofline_gemm']['achieved'],'cpu',d['cpu_baseline']['value'])
for x in d.get('extra_configs',[]): print('  extra',x['workload'][:50],x['value'],x['roofline']['frac'],x.get('roofline_gemm',{}).get('achieved'))"
  tail -5 gpurun_out/s21_bench.err
} > gpurun_out/s21.log 2>&1
cat gpurun_out/s21.log
