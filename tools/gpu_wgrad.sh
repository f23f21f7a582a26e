#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_layer_ops_gpu.py -q --timeout 120 2>&1 | tail -3
grep -E "wgrad" <(timeout 300 python tools/sweep_layers.py 2>&1) | tee gpurun_out/sweep_wgrad_pair.txt
grep -E "wgrad" <(SG_B200_NO_WGRAD_PAIR=1 timeout 300 python tools/sweep_layers.py 2>&1) | tee gpurun_out/sweep_wgrad_nopair.txt
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
print('headline ms', d['ms_per_step'], 'launches', d['gpu_launches_per_step'], 'roof', d['roofline']['frac'])
print({k: round(v['ms'], 3) for k, v in d['sdfnet'].items() if isinstance(v, dict) and 'ms' in v})
PY
SG_B200_NO_WGRAD_PAIR=1 timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_quick2.json 2>/dev/null; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_quick2.json'))
print('NO_WGRAD_PAIR headline ms', d['ms_per_step'])
print({k: round(v['ms'], 3) for k, v in d['sdfnet'].items() if isinstance(v, dict) and 'ms' in v})
PY
