#!/bin/bash
# (round 2, historical) tail split of the single-pass kernel: FLMIP_TAIL_SPLIT only exists with profiles/r2/08_tail_split_experiment.patch applied
# (measured and not adopted: profiles/r2/08_tail_split_ab.txt); parity first, then FLMIP_TAIL_SPLIT (split units per 100 resident CTAs) over the workloads
mkdir -p gpurun_out/r2split
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "tail_split or single_pass or slot_hand_off or baseline_configs or c3_full" 2>&1 | tail -5
run() { # name, workload, env...
  name=$1; w=$2; shift 2
  extra=""; [ $w = c3 ] && extra="--layers 256"
  env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered $extra > gpurun_out/r2split/${name}_$w.json 2> gpurun_out/r2split/${name}_$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2split/${name}_$w.json').read().strip().splitlines()[-1])
print('$name', '$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity_check']['mismatches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" 2>&1 | tail -1
}
for rep in 1 2; do
for w in c5 c3 c2; do
for sp in 0 50 100 200 400; do
  run s$sp $w FLMIP_TAIL_SPLIT=$sp
done
done
done
