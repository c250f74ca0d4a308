#!/bin/bash
# round-2 GPU session 7: full GPU suite (new tests), default bench with the tf32 extra, launch list of one epoch for profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
  echo "== full gpu suite"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -25
  echo "== default bench"; SECONDS=0; timeout 900 python bench.py > gpurun_out/bench_r2_s7.json 2> gpurun_out/bench_r2_s7.err; echo "rc=$? wall=${SECONDS}s"; grep -E "extra config|value arm|e2e arm" gpurun_out/bench_r2_s7.err
  python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_s7.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
print('gemm',d['roofline_gemm']['frac'],d['roofline_gemm']['achieved'])
for e in d['extra_configs']: print({k:e.get(k) for k in ('workload','value','ms_per_step','error')}, e.get('roofline_gemm'))
PY
} > gpurun_out/s7.log 2>&1
PHC_GRAPH_ROLLOUT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-points --no-extras > gpurun_out/s7_ncu.log 2>&1
gzip -f gpurun_out/launches_r2.csv
ls -la gpurun_out | tail -5
cat gpurun_out/s7.log
