#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k "halo" --timeout 120 2>&1 | tail -2
timeout 300 python tools/sweep_layers.py > gpurun_out/sweep.txt 2>&1; grep -E "conv 64|convT 128" gpurun_out/sweep.txt
SG_B200_NO_PAIR=1 timeout 300 python tools/sweep_layers.py > gpurun_out/sweep_nopair.txt 2>&1; grep -E "conv 64|convT 128" gpurun_out/sweep_nopair.txt
