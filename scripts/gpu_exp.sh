#!/bin/bash
# usage: gpu_exp.sh "<dbg flags list>" "<workloads>" [test]
set -u
mkdir -p gpurun_out; : > gpurun_out/exp.txt
if [ "${3:-}" = test ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/exp.txt; fi
for dbg in $1; do
  for w in $2; do
    extra=""
    [ $w = c3 ] && extra="--layers 256"
    [ $w = c4 ] && extra="--layers 4"
    r=$(timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $extra 2>&1 | python -c "
import sys,json
try:
  d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
")
    echo "${FLMIP_CTAS_PER_SM:-} ${FLMIP_STAGES:-} $w $r" | tee -a gpurun_out/exp.txt
  done
done
