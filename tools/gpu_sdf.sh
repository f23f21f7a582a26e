#!/bin/bash
# GPU visit focused on the fused SDFNet kernels: parity tests, timings, launch list, ncu full captures -> gpurun_out/
set -x
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q -k "sdfnet or autodecoder or hybrid" > gpurun_out/pytest_sdf.log 2>&1; RC=$?; echo "pytest-sdf rc=$RC $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
tail -25 gpurun_out/pytest_sdf.log
if [ $RC -ne 0 ]; then exit 1; fi
timeout 200 python tools/prof_sdf_fwd.py 250000 20 > gpurun_out/sdf_fwd.log 2>&1
timeout 200 python tools/prof_sdf_fwd.py 8388608 5 >> gpurun_out/sdf_fwd.log 2>&1
cat gpurun_out/sdf_fwd.log
timeout 300 python bench.py --workload autodecoder > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; tail -3 gpurun_out/bench_ad.err; cat gpurun_out/bench_ad.json
if [ "$1" != "quick" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ad.csv python bench.py --workload autodecoder --ad-shapes 64 --steps 2 --warmup 3 > gpurun_out/ncu_ad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_sdfnet_fwd -s 2 -c 1 -f -o gpurun_out/prof_sdf python tools/prof_sdf_fwd.py 1048576 2 > gpurun_out/ncu_sdf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_sdfnet_bwd -s 2 -c 1 -f -o gpurun_out/prof_sdf_bwd python bench.py --workload autodecoder --ad-shapes 64 --steps 1 --warmup 3 > gpurun_out/ncu_sdf_bwd.log 2>&1
fi
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
