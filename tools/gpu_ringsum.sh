#!/bin/bash
# ring-sum self-test per variant, then chain vs scan ring sums on several slab sizes: gpu_ringsum.sh <tag> lib1.so ...
TAG=$1; shift
mkdir -p gpurun_out
run() { # name lib mode nrad
  FARGO_B200_LIB=$PWD/$2 FARGO_B200_RINGSUM=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --nrad $4 --naz 16384 > gpurun_out/${TAG}_$1_$3_$4.log 2>&1
  tail -1 gpurun_out/${TAG}_$1_$3_$4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms_per_step']; print('$1 $3 $4', round(d['ms_per_step'],3), {n.replace('k_ring_mean',''):v for n,v in k.items() if 'ring_mean' in n or 'radial' in n})"
}
for lib in "$@"; do
  n=$(basename $lib .so)
  FARGO_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_math.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest_$n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_$n.log
  echo "$n: $(tail -2 gpurun_out/${TAG}_pytest_$n.log | tr '\n' ' ')"
  for nrad in 8192 2048 1038; do run $n $lib scan $nrad; done
done
for nrad in 8192 2048 1038; do run chain $1 chain $nrad; done
