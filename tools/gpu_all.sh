#!/bin/bash
# full GPU visit: parity suite, smoke, igemm diagnostics, SDFNet timings, bench lines
set -x
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; RC=$?; echo "pytest rc=$RC $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
tail -15 gpurun_out/pytest_gpu.log
if [ $RC -ne 0 ]; then exit 1; fi
timeout 300 python tools/diag_conv.py > gpurun_out/diag_conv.txt 2>&1; grep -E "==|diag= 0|diag=15|diag= 3|diag= 4" gpurun_out/diag_conv.txt
timeout 200 python tools/prof_sdf_fwd.py 250000 20 > gpurun_out/sdf_fwd.log 2>&1
timeout 200 python tools/prof_sdf_fwd.py 8388608 5 >> gpurun_out/sdf_fwd.log 2>&1
cat gpurun_out/sdf_fwd.log
timeout 300 python bench.py --workload autodecoder > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; cat gpurun_out/bench_ad.json
timeout 600 python bench.py > gpurun_out/bench_wgan.json 2> gpurun_out/bench_wgan.err; tail -3 gpurun_out/bench_wgan.err; cat gpurun_out/bench_wgan.json
timeout 400 python tools/sweep_layers.py > gpurun_out/sweep.txt 2>&1
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
