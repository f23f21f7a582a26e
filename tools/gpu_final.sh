#!/bin/bash
# end-of-round visit: parity suite, smoke, default bench line, launch list + full captures for profiles/
set -x
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; RC=$?; echo "pytest rc=$RC $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
tail -4 gpurun_out/pytest_gpu.log
if [ $RC -ne 0 ]; then tail -40 gpurun_out/pytest_gpu.log; exit 1; fi
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_wgan.json 2> gpurun_out/bench_wgan.err; tail -2 gpurun_out/bench_wgan.err; cat gpurun_out/bench_wgan.json
if [ "$1" == "prof" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wgan.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sdfnet > gpurun_out/ncu_wgan.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_igemm -s 3 -c 1 -f -o gpurun_out/prof_conv python tools/prof_conv.py 5 > gpurun_out/ncu_conv.log 2>&1
fi
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
