#!/bin/bash
# tuning sweep of the persistent single-pass kernel: CTAs per SM x ring depth, per workload
set -u
mkdir -p gpurun_out


: > gpurun_out/sweep.txt
for cfg in "2 3" "2 2" "2 1" "1 6" "1 4"; do
  set -- $cfg
  for w in c2 c5 c3 c4; do
    extra=""
    [ $w = c3 ] && extra="--layers 256"
    [ $w = c4 ] && extra="--layers 4"
    r=$(FLMIP_CTAS_PER_SM=$1 FLMIP_STAGES=$2 timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $extra 2>&1 | python -c "
import sys,json
try:
  d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
")
    echo "ctas=$1 stages=$2 $w $r" | tee -a gpurun_out/sweep.txt
  done
done
