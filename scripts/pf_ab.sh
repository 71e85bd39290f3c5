#!/bin/bash
# A/B of the scheduler prefetch depth (tuning builds libfloor_b200_mip_pf<N>.so) on the headline workloads
for w in c2 c5 c3 c1; do
  extra=""; [ $w = c3 ] && extra="--layers 256"
  for v in "" _pf1 _pf2 _pf3 "" _pf2; do
    r=$(FLMIP_LIB=/root/repo/floor_b200/libfloor_b200_mip$v.so timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $extra 2>&1 | python -c "
import sys,json
try:
  d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
")
    echo "$w [${v:-pf4}] $r"
  done
done
