#!/bin/bash
mkdir -p gpurun_out
export SG_B200_NO_REBUILD=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_final_check.json 2>gpurun_out/bench_final_check.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_final_check.json'))
print('headline ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'launches', d['gpu_launches_per_step'], 'roof', round(d['roofline']['frac'], 4), d['roofline'].get('same_kernel_other_conditions'))
print({k: round(v.get('ms_per_step', v.get('ms_per_batch_5to1_schedule', 0)), 3) for k, v in d.get('configs', {}).items() if isinstance(v, dict)})
PY
