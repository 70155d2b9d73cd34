#!/bin/bash
# round 2, tenth GPU pass: tile length of the scan ring sum (RS_TL 8 / 16 / 32) on a slab of one of 8 GPUs (1052 rings) and of 4 (2076)
TAG=${1:-r02_v15}
mkdir -p gpurun_out
for lib in "" exp/rs_tl16.so exp/rs_tl32.so; do
  for nrad in 1052 2076; do
    name=$(basename "${lib:-base}" .so)
    FARGO_B200_LIB=${lib:+$PWD/$lib} FARGO_B200_RINGSUM=scan timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-tolerance-mode --nrad $nrad > gpurun_out/${TAG}_${name}_$nrad.log 2>&1
    grep '^{' gpurun_out/${TAG}_${name}_$nrad.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['kernels_ms_per_step']
print('$name', $nrad, 'rings:', round(d['ms_per_step'], 4), 'ms/step', d['checksum']['sha256'][:10], {x: k[x] for x in k if 'ring_mean' in x})"
  done
  FARGO_B200_LIB=${lib:+$PWD/$lib} timeout 300 python -m pytest tests/test_gpu_math.py -m gpu -q 2>&1 | tail -1
done
