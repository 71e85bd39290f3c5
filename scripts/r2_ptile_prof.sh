#!/bin/bash
# ncu --set full capture of the persistent tile kernel: scripts/r2_ptile_prof.sh <workload> <tag> [kernel regex]
set -u
w=$1; tag=$2; rx=${3:-flmip_ptile}
mkdir -p gpurun_out/r2prof
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -f -o gpurun_out/r2prof/prof_${tag} python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-incumbent --no-layered > gpurun_out/r2prof/ncu_full_${tag}.log 2>&1
tail -3 gpurun_out/r2prof/ncu_full_${tag}.log
python scripts/ncu_summary.py gpurun_out/r2prof/prof_${tag}.ncu-rep > gpurun_out/r2prof/summary_${tag}.txt 2>&1
cat gpurun_out/r2prof/summary_${tag}.txt
