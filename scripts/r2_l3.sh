#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "persistent" 2>&1 | tail -4
bash scripts/r2_ab.sh "n2" l3
FLMIP_PTILE_STREAM_L3=0 bash scripts/r2_ab.sh "n2" l3
