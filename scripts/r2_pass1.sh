#!/bin/bash
# round 2, first GPU pass: whole GPU suite, smoke, default bench line (with layered legs + parity_check), reference arm
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a/box.txt; nproc >> gpurun_out/r2a/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r2a/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_c2.json 2> gpurun_out/r2a/bench_c2.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2a/bench_ref.json 2>&1
tail -5 gpurun_out/r2a/pytest_gpu.log; cat gpurun_out/r2a/smoke.log | tail -2; cat gpurun_out/r2a/bench_c2.json | cut -c1-3000; tail -3 gpurun_out/r2a/bench_c2.err
