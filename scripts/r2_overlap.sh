#!/bin/bash
# chain overlap (flmip_stream_set_chain_overlap): its tests, the whole GPU suite, and the pipelined leg of the bench per workload
mkdir -p gpurun_out/r2ov
timeout 900 python -m pytest tests/test_gpu_overlap.py -x -q -m gpu -s 2>&1 | tail -8
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in c2 c1 c5 n1 n2; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered > gpurun_out/r2ov/$w.json 2> gpurun_out/r2ov/$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2ov/$w.json').read().strip().splitlines()[-1])
p=d['pipelined']
print('$w', 'strict', d['value'], d['ms_per_step'], d['roofline']['frac'], '| pipelined', p['value'], p['ms_per_step'], p['frac'], '| mismatches', d['parity_check']['mismatches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
" 2>&1 | tail -1
done
