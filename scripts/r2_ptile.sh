#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "persistent or stress" 2>&1 | tail -15 > gpurun_out/r2c/pytest_ptile.log; tail -5 gpurun_out/r2c/pytest_ptile.log
for w in n1 n2 c2 c3 c1 c5; do
  extra=""; [ $w = c3 ] && extra="--layers 256"
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered $extra > gpurun_out/r2c/bench_$w.json 2> gpurun_out/r2c/bench_$w.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2c/bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['detail'], d['parity_check']['mismatches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" 2>&1 | tail -2; tail -2 gpurun_out/r2c/bench_$w.err
done
