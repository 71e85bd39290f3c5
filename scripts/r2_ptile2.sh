#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "persistent" 2>&1 | tail -15 > gpurun_out/r2c/pytest_ptile.log; tail -5 gpurun_out/r2c/pytest_ptile.log
bash scripts/r2_ab.sh "n2 n1" u
