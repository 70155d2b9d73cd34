#!/bin/bash
# ncu --set full of ONE kernel (regex) on the 4096x8192 grid: gpu_ncu_kernel.sh <tag> <kernel-regex> [skip]
TAG=$1; RE=$2; SKIP=${3:-3}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nrad 4096 --naz 8192 > gpurun_out/${TAG}.log 2>&1
tail -2 gpurun_out/${TAG}.log | cut -c1-300
