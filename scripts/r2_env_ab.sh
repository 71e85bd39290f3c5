#!/bin/bash
mkdir -p gpurun_out/r2env
run() { # name, env..., workload
  name=$1; w=$2; shift 2
  extra=""; [ $w = c3 ] && extra="--layers 256"
  for rep in 1 2; do
  env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered $extra > gpurun_out/r2env/${name}_$w.json 2> gpurun_out/r2env/${name}_$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2env/${name}_$w.json').read().strip().splitlines()[-1])
print('$name', '$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_check']['mismatches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" 2>&1 | tail -1
  done
}
run base c5 X=1
run st3 c5 FLMIP_STAGES=3
run base c3 X=1
run st3 c3 FLMIP_STAGES=3
run base c2 X=1
run st3 c2 FLMIP_STAGES=3
