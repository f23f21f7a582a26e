#!/bin/bash
mkdir -p gpurun_out
export SG_B200_NO_REBUILD=1
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_parity_gpu.py tests/test_steps_gpu.py tests/test_layer_ops_gpu.py -q --timeout 600 -x 2>&1 | tail -4
q() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], 'headline ms', round(d['ms_per_step'], 4), 'launches', d['gpu_launches_per_step'], {k: round(v['ms_per_step'], 3) for k, v in d.get('configs', {}).items() if isinstance(v, dict) and 'ms_per_step' in v})
PY
}
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bq_base.json 2>/dev/null; q gpurun_out/bq_base.json
SG_B200_B_TMA=1 timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bq_btma.json 2>/dev/null; q gpurun_out/bq_btma.json
echo "---- B_TMA=1"; SG_B200_B_TMA=1 timeout 300 python tools/sweep_layers.py 2>&1 | grep -E "^B=( 64|192).*(convT 128|conv 64)" | grep -v wgrad
echo "---- default"; timeout 300 python tools/sweep_layers.py 2>&1 | grep -E "^B=( 64|192).*(convT 128|conv 64)" | grep -v wgrad
