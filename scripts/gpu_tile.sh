#!/bin/bash
# tile-kernel pass: parity tests, then the NPOT bench workloads (n1, n2) with the tile kernel vs the literal per-level kernel
set -u
tag=${1:-tile}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -14 $out/pytest_gpu.log
for w in n1 n2; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_$w.json 2> $out/bench_$w.err; cut -c1-330 $out/bench_$w.json
done
