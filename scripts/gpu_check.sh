#!/bin/bash
# Quick GPU pass: parity tests, smoke, the default bench line and the reference arm.  usage: gpu_check.sh <tag> [workloads...]
set -u
tag=${1:-check}; shift || true
out=gpurun_out/$tag
mkdir -p $out
nproc > $out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -14 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2>&1; cut -c1-300 $out/bench_ref.json
for w in ${@:-c2}; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $out/bench_$w.json 2> $out/bench_$w.err; cut -c1-600 $out/bench_$w.json
done
