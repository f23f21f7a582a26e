#!/bin/bash
mkdir -p gpurun_out
export SG_B200_NO_REBUILD=1
timeout 300 python tools/exp_graph_bubbles.py 2>&1 | tee gpurun_out/exp_graph_bubbles.txt | tail -8
