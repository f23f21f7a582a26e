#!/bin/bash
mkdir -p gpurun_out
export SG_B200_NO_REBUILD=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6
echo "---- new"; timeout 300 python tools/sweep_layers.py 2>&1 | grep -E "^B=( 64|128|192)" | tee gpurun_out/sweep_ab4_new.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
print('headline ms', d['ms_per_step'], 'launches', d['gpu_launches_per_step'], 'roof', d['roofline']['frac'])
print({k: round(v['ms'], 3) for k, v in d['sdfnet'].items() if isinstance(v, dict) and 'ms' in v})
print({k: round(v['ms_per_step'], 3) for k, v in d.get('configs', {}).items() if isinstance(v, dict) and 'ms_per_step' in v})
PY
