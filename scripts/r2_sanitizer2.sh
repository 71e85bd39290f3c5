#!/bin/bash
out=gpurun_out/r2_sanitizer2.txt
echo "compute-sanitizer pass of round 2, second part (chain overlap: late-waiting kernels; cp.async gather of the group / layer patches), B200" > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool python scripts/race_overlap.py" >> $out
  timeout 900 compute-sanitizer --tool $tool python scripts/race_overlap.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========$" | cut -c1-330 | tail -40 >> $out
done
cat $out | tail -70
