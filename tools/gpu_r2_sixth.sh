#!/bin/bash
# round 2, sixth GPU pass: AlphaMode 1 on the device, the two-pass CFL reduction (screen + candidates) against the full one
TAG=${1:-r02_v10}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_screen.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench_screen.log
tail -2 gpurun_out/${TAG}_bench_screen.log | cut -c1-3000
FARGO_B200_CFL=full timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_full.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench_full.log
tail -2 gpurun_out/${TAG}_bench_full.log | cut -c1-3000
