#!/bin/bash
# Full GPU pass for the record: parity tests, smoke, bench lines, ncu launch lists + full captures.  usage: gpu_pass.sh <tag>
set -u
tag=${1:-pass}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
nproc > $out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -1 $out/smoke.log
for w in c2 c1 c5 c3 c4 n1 n2; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $out/bench_$w.json 2> $out/bench_$w.err; cut -c1-400 $out/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2>&1
for w in c2 c5 c3; do
  extra=""; [ $w = c3 ] && extra="--layers 256"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline $extra > $out/ncu_launch_$w.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:flmip_fast -s 3 -c 1 -f -o $out/prof_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $extra > $out/ncu_full_$w.log 2>&1
  # gpurun_out/ is limited to 64 MiB: summarise on the box, keep only the C2 report itself
  python scripts/ncu_summary.py $out/prof_$w.ncu-rep > $out/ncu_full_summary_$w.txt 2>&1
  [ $w = c2 ] || rm -f $out/prof_$w.ncu-rep
done
ls -la $out
