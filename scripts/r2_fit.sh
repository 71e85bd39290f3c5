#!/bin/bash
# fixed cost of one chain, final kernels of round 2: C3 at growing layer counts (t = t0 + layers * t1), C2-shaped images of growing size
set -u
mkdir -p gpurun_out; : > gpurun_out/fit2.txt
for L in 8 32 64 128 256 512 1024 2048; do
  r=$(timeout 300 python bench.py --workload c3 --layers $L --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-incumbent --no-layered 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])")
  echo "c3 layers=$L $r" | tee -a gpurun_out/fit2.txt
done
python - <<'PY' | tee -a gpurun_out/fit2.txt
import numpy as np
rows=[l.split() for l in open('gpurun_out/fit2.txt') if l.startswith('c3')]
L=np.array([int(r[1].split('=')[1]) for r in rows],float); t=np.array([float(r[3]) for r in rows])*1e3
A=np.vstack([np.ones_like(L),L]).T
(t0,t1),*_=np.linalg.lstsq(A,t,rcond=None)
print(f"fit over all: t = {t0:.2f} us + layers * {t1:.4f} us   (5 592 404 B per layer -> {5592404/t1/1e3:.0f} GB/s asymptotic)")
m=L>=128
(t0,t1),*_=np.linalg.lstsq(A[m],t[m],rcond=None)
print(f"fit over layers >= 128: t = {t0:.2f} us + layers * {t1:.4f} us   ({5592404/t1/1e3:.0f} GB/s asymptotic)")
PY
