#!/bin/bash
# one ncu --set full capture of every kernel of one hydro step (C4-size grid, larger than L2)
# usage: gpu_ncu_full.sh <tag> [nrad naz]
TAG=${1:-prof}; NRAD=${2:-4096}; NAZ=${3:-8192}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s 80 -c 28 -f -o gpurun_out/${TAG} \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nrad $NRAD --naz $NAZ > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
ls -la gpurun_out/
