#!/bin/bash
# One GPU pass: parity tests, bench lines for the single-GPU workloads, ncu launch list + full capture of the top kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for w in c2 c1 c5 c3; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 1500 gpurun_out/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1
for w in c2 c5 c3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$w.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:flmip_fast -s 3 -c 2 -f -o gpurun_out/prof_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$w.log 2>&1
done
ls -la gpurun_out
