#!/bin/bash
# builds tuning variants of the library (dev subset of kernels) into build/variants/: scripts/r2_variants.sh name "DEFS" ...
set -e
mkdir -p build/variants
cd floor_b200/csrc
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  make -s -B DEFS="-DFLMIP_DEV_ONLY $defs" CUBIN=/tmp/var_$name.cubin OUT=../../build/variants/lib_$name.so >/dev/null
  echo built $name
done
