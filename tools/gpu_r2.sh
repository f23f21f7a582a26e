#!/bin/bash
# round-2 GPU visit: parity suite (errors recorded), smoke, bench line, layer sweep, launch list.  Args: [record] [prof]
mkdir -p gpurun_out
T0=$(date +%s)
if [[ " $* " == *" record "* ]]; then export SG_PARITY_RECORD=1; fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; RC=$?
unset SG_PARITY_RECORD
echo "pytest rc=$RC $(( $(date +%s)-T0 ))s" | tee gpurun_out/times.log
tail -5 gpurun_out/pytest_gpu.log
if [ $RC -ne 0 ]; then grep -E "^(E  |FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -40; fi
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
tail -3 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; echo
timeout 300 python tools/sweep_layers.py $SWEEP_ARGS > gpurun_out/sweep.txt 2>&1; cat gpurun_out/sweep.txt
if [[ " $* " == *" prof "* ]]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wgan_gp.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sdfnet --no-extra > gpurun_out/ncu_wgan.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wgan.csv python bench.py --workload wgan --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-sdfnet --no-extra > gpurun_out/ncu_wgan2.log 2>&1
fi
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
