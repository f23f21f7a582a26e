#!/bin/bash
mkdir -p gpurun_out
export SG_B200_NO_REBUILD=1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -4
timeout 600 python tools/prof_step.py wgan_gp > gpurun_out/step_kernels_wgan_gp.txt 2> gpurun_out/step_kernels.err; head -12 gpurun_out/step_kernels_wgan_gp.txt | cut -c1-110
q() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], 'headline ms', round(d['ms_per_step'], 4), 'launches', d['gpu_launches_per_step'], {k: round(v['ms_per_step'], 3) for k, v in d.get('configs', {}).items() if isinstance(v, dict) and 'ms_per_step' in v})
PY
}
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bq_base.json 2>/dev/null; q gpurun_out/bq_base.json
