#!/bin/bash
mkdir -p gpurun_out/r2b
for w in c1 c2 c5 c3 n1; do
  timeout 600 python bench.py --impl incumbent --workload $w --steps 5 --warmup 3 > gpurun_out/r2b/incumbent_$w.json 2> gpurun_out/r2b/incumbent_$w.err
  cut -c1-1500 gpurun_out/r2b/incumbent_$w.json | sed 's/.*gpu_incumbent/gpu_incumbent/'; tail -3 gpurun_out/r2b/incumbent_$w.err
done
