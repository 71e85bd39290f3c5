#!/bin/bash
# (round 2, historical) first late-wait experiment: FLMIP_LATE_WAIT_EXPERIMENT only exists with profiles/r2/08_tail_split_experiment.patch applied;
# the adopted form is flmip_stream_set_chain_overlap (scripts/r2_overlap.sh, bench.py key "pipelined")
mkdir -p gpurun_out/r2lw
export FLMIP_LIB=$PWD/build/variants/lib_lw.so
run() { # name, workload, env...
  name=$1; w=$2; shift 2
  extra=""; [ $w = c3 ] && extra="--layers 256"
  env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered $extra > gpurun_out/r2lw/${name}_$w.json 2> gpurun_out/r2lw/${name}_$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2lw/${name}_$w.json').read().strip().splitlines()[-1])
print('$name', '$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_check']['mismatches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" 2>&1 | tail -1
}
for rep in 1 2; do
for w in c1 c5 c3 c2; do
for lw in 0 1; do
  run lw$lw $w FLMIP_LATE_WAIT_EXPERIMENT=$lw
done
done
done
