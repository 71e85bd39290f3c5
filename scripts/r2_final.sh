#!/bin/bash
# Round-2 GPU pass for the record: parity tests, smoke, bench lines (default line with layered legs, parity_check, PCIe ceiling,
# gpu_incumbent), per-workload lines, ncu launch lists + full captures.  usage: r2_final.sh <tag>
set -u
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/box.txt 2>&1; nproc >> $out/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; tail -1 $out/smoke.log
t0=$SECONDS
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_default.json 2> $out/bench_default.err; cut -c1-300 $out/bench_default.json
echo "default bench.py line: $((SECONDS - t0)) s wall" | tee $out/bench_default.time
for w in c1 c5 c3 c4 n1 n2; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $out/bench_$w.json 2> $out/bench_$w.err; cut -c1-200 $out/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/bench_ref.json 2>&1
timeout 600 python bench.py --impl incumbent --steps 10 --warmup 3 > $out/bench_incumbent.json 2>&1
for w in c2 c5 c3 n2; do
  extra=""; [ $w = c3 ] && extra="--layers 256"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-incumbent --no-layered $extra > $out/ncu_launch_$w.log 2>&1
  rx=flmip_fast; [ $w = n2 ] && rx=flmip_ptile
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o $out/prof_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-incumbent --no-layered $extra > $out/ncu_full_$w.log 2>&1
  python scripts/ncu_summary.py $out/prof_$w.ncu-rep > $out/ncu_full_summary_$w.txt 2>&1
  [ $w = n2 ] || rm -f $out/prof_$w.ncu-rep
done
ls -la $out
