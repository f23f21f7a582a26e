#!/bin/bash
# profiling visit: launch list of the WGAN step (graph-less) + full captures of the roofline-probe conv kernel
set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_wgan.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_wgan.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_igemm -s 3 -c 1 -f -o gpurun_out/prof_conv python tools/prof_conv.py 5 > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sg_sdfnet_fwd -s 2 -c 1 -f -o gpurun_out/prof_sdf python tools/prof_sdf_fwd.py 1048576 2 > gpurun_out/ncu_sdf.log 2>&1
echo done
