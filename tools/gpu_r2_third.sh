#!/bin/bash
# round 2, third GPU pass: cp.async ring staging in all marching kernels + marching CFL kernel: parity suite, bench line, variants, ncu --set full
TAG=${1:-r02_v3}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-4500
STEPS=10 BENCH_ARGS="--no-tolerance-mode" bash tools/gpu_variants.sh ${TAG}_var $VARIANTS
grep -h -o '"sha256": "[0-9a-f]*"' gpurun_out/${TAG}_var_*.log | sort | uniq -c
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 60 -c 12 -f -o gpurun_out/${TAG}_full_c5 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_full_c5.log 2>&1
tail -2 gpurun_out/${TAG}_full_c5.log | cut -c1-300
ls -la gpurun_out/ | tail -14
