#!/bin/bash
# ncu --set full capture of the tile kernel: scripts/gpu_tile_prof.sh <workload> <tag>
set -u
w=$1; tag=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flmip_tile -s 6 -c 1 -f -o gpurun_out/prof_${tag} python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_full_${tag}.log 2>&1
tail -3 gpurun_out/ncu_full_${tag}.log
