#!/bin/bash
# A/B of library variants built by scripts/r2_variants.sh: scripts/r2_ab.sh "<workloads>" <variant names...>
wls=$1; shift
mkdir -p gpurun_out/r2ab
for rep in 1 2; do
for v in "$@"; do
  for w in $wls; do
    extra=""; [ $w = c3 ] && extra="--layers 256"
    FLMIP_LIB=$PWD/build/variants/lib_$v.so timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered $extra > gpurun_out/r2ab/${v}_$w.json 2> gpurun_out/r2ab/${v}_$w.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/r2ab/${v}_$w.json').read().strip().splitlines()[-1])
print('$v', '$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_check']['mismatches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" 2>&1 | tail -1; tail -1 gpurun_out/r2ab/${v}_$w.err
  done
done
done
