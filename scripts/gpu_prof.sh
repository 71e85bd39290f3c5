#!/bin/bash
# ncu --set full capture of the single-pass kernel for one workload: scripts/gpu_prof.sh <workload> <tag> [bench args...]
set -u
w=$1; tag=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flmip_fast -s 3 -c 1 -f -o gpurun_out/prof_${tag} python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_full_${tag}.log 2>&1
tail -3 gpurun_out/ncu_full_${tag}.log
