#!/bin/bash
# end-of-round single-GPU visit: whole suite (gates enforced), smoke, default bench line + reference arm, launch lists + ncu captures for profiles/
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; RC=$?
echo "pytest rc=$RC $(( $(date +%s)-T0 ))s" | tee gpurun_out/times.log
tail -4 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
tail -3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
    print('headline ms', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches_per_step'], 'roof', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['config']['step_frac_of_sustained_peak'])
    for k, v in d.get('configs', {}).items():
        print(' ', k, v.get('ms_per_step', v.get('ms_per_batch_5to1_schedule')), v.get('error', ''), v.get('gpu_launches_per_step', ''), v.get('step_frac_of_sustained_peak', ''))
        for sub in ('d_update', 'g_update'):
            if sub in v: print('     ', sub, v[sub]['ms_per_step'], v[sub]['step_frac_of_sustained_peak'])
    s = d.get('sdfnet', {})
    print('  sdfnet', {k: (round(v['ms'], 3), round(v.get('frac_of_burst_peak', v.get('frac_of_sustained_peak', 0)), 3)) for k, v in s.items() if isinstance(v, dict) and 'ms' in v})
    print('  cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY
if [[ " $* " == *" ref "* ]]; then
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$? $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log; head -c 400 gpurun_out/bench_ref.json; echo
fi
if [[ " $* " == *" prof "* ]]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wgan_gp.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sdfnet --no-extra > gpurun_out/ncu_wgan.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ad.csv python bench.py --workload autodecoder --ad-shapes 64 --steps 2 --warmup 3 > gpurun_out/ncu_ad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_igemm -s 3 -c 1 -f -o gpurun_out/prof_conv python tools/prof_conv.py 5 > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_sdfnet_fwd -s 2 -c 1 -f -o gpurun_out/prof_sdf python tools/prof_sdf_fwd.py > gpurun_out/ncu_sdf.log 2>&1
timeout 300 python tools/sweep_layers.py > gpurun_out/sweep.txt 2>&1
timeout 300 python tools/prof_step.py wgan_gp > gpurun_out/step_kernels_wgan_gp.txt 2>/dev/null
timeout 300 python tools/prof_step.py gan > gpurun_out/step_kernels_gan.txt 2>/dev/null
for tool in racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_igemm.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/times.log; tail -4 gpurun_out/sanitizer_$tool.log
done
# clock traces need the instrumented build: last, the box is discarded afterwards
timeout 600 python -m shapegan_b200.build --force --trace > /dev/null 2>&1
SG_B200_NO_REBUILD=1 timeout 300 python tools/trace_igemm.py > gpurun_out/trace_conv.txt 2>&1
SG_B200_NO_REBUILD=1 timeout 300 python tools/trace_igemm.py convt > gpurun_out/trace_convt.txt 2>&1
SG_B200_NO_REBUILD=1 timeout 300 python tools/trace_igemm.py patch > gpurun_out/trace_patch.txt 2>&1
fi
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
