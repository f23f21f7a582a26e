#!/bin/bash
export SG_B200_NO_REBUILD=1     # the snapshot carries the library built in the dev container; never race nvcc across ranks
# multi-GPU visit (gpurun --gpus N): the fused peer-memory optimizer step vs the NCCL path.  Args: N [notest]
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L | head -8
if [ "$2" != "notest" ]; then
timeout 600 python -m pytest tests/test_dp_gpu.py -q --timeout 300 > gpurun_out/pytest_dp.log 2>&1; echo "dp tests rc=$? $(( $(date +%s)-T0 ))s" | tee gpurun_out/times_multi.log
tail -5 gpurun_out/pytest_dp.log; grep -E "^E  " gpurun_out/pytest_dp.log | head -20
fi
for MODE in ${MODES:-fused nccl}; do
  EXTRA=""; if [ "$MODE" == "nccl" ]; then EXTRA="--no-extra"; fi
  SG_B200_DP=$MODE timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-sdfnet $EXTRA \
     > gpurun_out/bench_n${N}_${MODE}.json 2> gpurun_out/bench_n${N}_${MODE}.err
  echo "bench N=$N $MODE rc=$? $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times_multi.log
  grep -v "NCCL INFO" gpurun_out/bench_n${N}_${MODE}.err | grep -v "^\*\|OMP_NUM\|^$" | tail -5
  echo "NCCL INFO lines: $(grep -c 'NCCL INFO' gpurun_out/bench_n${N}_${MODE}.err); nranks: $(grep -o 'nranks [0-9]*' gpurun_out/bench_n${N}_${MODE}.err | sort | uniq -c | head -3)"
  wc -l gpurun_out/bench_n${N}_${MODE}.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_n${N}_${MODE}.json'))
    print('$MODE N=$N ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'dp', d['config'].get('data_parallel'))
    for k, v in d.get('configs', {}).items():
        print('   ', k, v.get('ms_per_step', v.get('ms_per_batch_5to1_schedule')), v.get('error', ''))
except Exception as e:
    print('parse failed', e)
PY
done
echo "done $(( $(date +%s)-T0 ))s" | tee -a gpurun_out/times_multi.log
