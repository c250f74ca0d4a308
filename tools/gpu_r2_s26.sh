#!/bin/bash
# round-2 GPU session 26: what the driver runs at round end, on the final commit: GPU suite, smoke(), default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== bench.py (default)"; timeout 1200 python bench.py 2> gpurun_out/s26_bench.err | tee gpurun_out/bench_r2_final_1gpu_b.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'] and d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['kernel_us'],[round(p['frac'],3) for p in d['roofline'].get('points',[])],'gemm',d['roofline_gemm']['achieved'],'cpu',d['cpu_baseline']['value'],'clocks',d['clocks'])
for x in d.get('extra_configs',[]): print('  extra',x['workload'][:50],x['value'],x['roofline']['frac'],x.get('roofline_gemm',{}).get('achieved'))"
  tail -3 gpurun_out/s26_bench.err
} > gpurun_out/s26.log 2>&1
cat gpurun_out/s26.log
