#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, sweeps, ncu launch lists + full captures -> gpurun_out/
set -x
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/times.log
timeout 600 python bench.py > gpurun_out/bench_wgan.json 2> gpurun_out/bench_wgan.err; echo "bench rc=$?" | tee -a gpurun_out/times.log
timeout 300 python bench.py --workload wgan_gp --no-cpu-baseline > gpurun_out/bench_wgan_gp.json 2> gpurun_out/bench_wgan_gp.err
timeout 300 python bench.py --workload autodecoder > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 200 python tools/prof_sdf_fwd.py 250000 20 > gpurun_out/sdf_fwd.log 2>&1
timeout 200 python tools/prof_sdf_fwd.py 8388608 5 >> gpurun_out/sdf_fwd.log 2>&1
timeout 400 python tools/sweep_layers.py > gpurun_out/sweep.txt 2>&1
echo "pre-ncu $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wgan.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_wgan.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ad.csv python bench.py --workload autodecoder --ad-shapes 64 --steps 2 --warmup 3 > gpurun_out/ncu_ad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_igemm -s 3 -c 2 -f -o gpurun_out/prof_conv python tools/prof_conv.py 5 > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_sdfnet_fwd -s 2 -c 1 -f -o gpurun_out/prof_sdf python tools/prof_sdf_fwd.py 1048576 2 > gpurun_out/ncu_sdf.log 2>&1
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
cat gpurun_out/bench_wgan.json gpurun_out/bench_ad.json gpurun_out/sdf_fwd.log
