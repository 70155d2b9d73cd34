#!/bin/bash
# round 2, seventh GPU pass: side-stream ring sums with small chunks + high priority (beside the radial sweep) against the earlier arrangement
TAG=${1:-r02_v11}
mkdir -p gpurun_out
for v in 32 0 64 16; do
FARGO_B200_RM_SIDE_CHUNK=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_side$v.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench_side$v.log
tail -2 gpurun_out/${TAG}_bench_side$v.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('side chunk $v:', d['ms_per_step'], d['checksum']['sha256'][:12], d['e2e']['value'], d['e2e'].get('at_one_snapshot_per_10K_steps', {}).get('value'), d['kernels_ms_per_step'])
"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
