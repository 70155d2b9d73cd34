#!/bin/bash
# parity tests on some variants, then time all variants: gpu_sweep.sh <tag> <parity1.so,parity2.so,...> lib1.so lib2.so ...
TAG=$1; PVS=$2; shift; shift
mkdir -p gpurun_out
for PV in ${PVS//,/ }; do
  n=$(basename $PV .so)
  FARGO_B200_LIB=$PWD/$PV timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_math.py -m gpu -x -q > gpurun_out/${TAG}_pytest_$n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_$n.log
  echo "parity $n: $(tail -3 gpurun_out/${TAG}_pytest_$n.log | tr '\n' ' ')"
done
bash tools/gpu_variants.sh $TAG "$@"
