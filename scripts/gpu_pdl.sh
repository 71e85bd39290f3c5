#!/bin/bash
# A/B of programmatic dependent launch (FLMIP_PDL=0 / 1) over workloads, then the parity suite with it on
set -u
for w in c2 c5 c1 n1 n2 c3; do
  extra=""; [ $w = c3 ] && extra="--layers 256"
  for v in 0 1 0 1; do
    r=$(FLMIP_PDL=$v timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $extra 2>&1 | python -c "
import sys,json
try:
  d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
")
    echo "$w PDL=$v $r"
  done
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
