#!/bin/bash
# round 2, eighth GPU pass: artificial-viscosity stage inside the source-term kernel (FARGO_B200_FUSE_ARTVISC=0: own kernel)
TAG=${1:-r02_v12}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_diagnostics.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for v in 1 0; do
FARGO_B200_FUSE_ARTVISC=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_av$v.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench_av$v.log
tail -2 gpurun_out/${TAG}_bench_av$v.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('fuse artvisc $v:', d['ms_per_step'], d['checksum']['sha256'][:12], d['e2e']['value'], d['kernels_ms_per_step'])
    elif 'rc=' not in l: print(l[:300])
"
done
