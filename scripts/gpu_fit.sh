#!/bin/bash
# fixed overhead of one launch: C3 at growing layer counts (t = t0 + layers * t1), and the NPOT workloads
set -u
mkdir -p gpurun_out; : > gpurun_out/fit.txt
for L in 8 32 64 128 256 512 1024 2048; do
  r=$(timeout 300 python bench.py --workload c3 --layers $L --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])")
  echo "c3 layers=$L $r" | tee -a gpurun_out/fit.txt
done
for w in n1 n2; do
  r=$(timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['launches_per_step'])")
  echo "$w $r" | tee -a gpurun_out/fit.txt
done
