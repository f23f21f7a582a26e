#!/bin/bash
export SG_B200_NO_REBUILD=1     # the snapshot carries the library built in the dev container; never race nvcc across ranks
# GPU visit: the whole -m gpu suite (parity errors recorded when "record" is passed) + smoke.  Args: [record] [pytest -k expression]
mkdir -p gpurun_out
T0=$(date +%s)
if [[ "$1" == "record" ]]; then export SG_PARITY_RECORD=1; shift; fi
KEXPR=()
if [[ -n "$1" ]]; then KEXPR=(-k "$1"); fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 "${KEXPR[@]}" > gpurun_out/pytest_gpu.log 2>&1; RC=$?
unset SG_PARITY_RECORD
echo "pytest rc=$RC $(( $(date +%s)-T0 ))s" | tee gpurun_out/times.log
tail -5 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -60
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times.log
