#!/bin/bash
out=gpurun_out/r2_sanitizer.txt
echo "compute-sanitizer pass of round 2 (persistent TMA tile kernel in both forms, literal kernel with 3-channel / packed formats), B200" > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool python scripts/race_ptile.py" >> $out
  timeout 900 compute-sanitizer --tool $tool python scripts/race_ptile.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========$" | cut -c1-330 | tail -40 >> $out
done
echo "== compute-sanitizer --tool memcheck pytest -k 'persistent or three_channel'" >> $out
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "persistent_tma_tile_kernel_all_formats or three_channel" 2>&1 | tail -4 >> $out
cat $out | tail -60
