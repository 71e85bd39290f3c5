#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "persistent or tile" 2>&1 | tail -15 > gpurun_out/r2c/pytest_ptile.log; tail -7 gpurun_out/r2c/pytest_ptile.log
bash scripts/r2_ab.sh "n2 n1" u
python scripts/ptile_sweep.py > gpurun_out/r2c/ptile_sweep.txt 2>&1; cat gpurun_out/r2c/ptile_sweep.txt
