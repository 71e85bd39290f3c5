#!/bin/bash
# multi-GPU pass on one box: bench.py under torchrun for N ranks.  usage: gpu_scale.sh <N> <tag> [workloads...]
set -u
n=$1; tag=$2; shift 2
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $out/gpus.txt 2>&1
for w in ${@:-c2 c3 c4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --steps 20 --warmup 5 > $out/bench_${w}_n$n.json 2> $out/bench_${w}_n$n.err
  grep '^{' $out/bench_${w}_n$n.json | cut -c1-330
done
