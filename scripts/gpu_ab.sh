#!/bin/bash
# A/B of tuning builds of the library: scripts/gpu_ab.sh "<lib paths (or 'default')>" "<workloads>" [test]
set -u
mkdir -p gpurun_out; : > gpurun_out/ab.txt
for lib in $1; do
  if [ "$lib" = default ]; then unset FLMIP_LIB; else export FLMIP_LIB=$PWD/$lib; fi
  if [ "${3:-}" = test ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab.txt; fi
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
    echo "$lib $w $r" | tee -a gpurun_out/ab.txt
  done
done
