#!/bin/bash
# round 2, GPU pass: parity suite, BASELINE configs on the product vs the reference, bench line (damping folded, tolerance mode), launch list
TAG=${1:-r02_v2}
mkdir -p gpurun_out
grep -m1 "model name" /proc/cpuinfo > gpurun_out/${TAG}_host.txt; nproc >> gpurun_out/${TAG}_host.txt; grep -o -w "avx512f\|avx2\|fma" /proc/cpuinfo | sort -u | tr '\n' ' ' >> gpurun_out/${TAG}_host.txt; df -h /tmp | tail -1 >> gpurun_out/${TAG}_host.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
timeout 1500 python tests/checkers/baseline_configs_on_gpu.py ${TAG} > gpurun_out/${TAG}_baseline_configs.log 2>&1; echo "baseline configs rc=$?" >> gpurun_out/${TAG}_baseline_configs.log
tail -12 gpurun_out/${TAG}_baseline_configs.log | cut -c1-500
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-4500
FARGO_B200_FOLD_DAMPING=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_nofold.log 2>&1
tail -1 gpurun_out/${TAG}_bench_nofold.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out/ | tail -12
# variants: the fused kernels with the old 4 + 4 halo columns against the default 2 + 2
STEPS=10 BENCH_ARGS="--no-tolerance-mode" bash tools/gpu_variants.sh ${TAG}_var exp/fs44.so
