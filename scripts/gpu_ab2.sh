#!/bin/bash
# A/B of two builds of the library: gpu_ab2.sh <suffix> "<workloads>"   (floor_b200/libfloor_b200_mip<suffix>.so vs the default)
set -u
suf=$1
for w in $2; do
  extra=""
  [ $w = c3 ] && extra="--layers 256"
  [ $w = c4 ] && extra="--layers 8"
  for v in "" $suf "" $suf; do
    r=$(FLMIP_LIB=/root/repo/floor_b200/libfloor_b200_mip$v.so timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $extra 2>&1 | python -c "
import sys,json
try:
  d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
")
    echo "$w [${v:-base}] $r"
  done
done
