#!/bin/bash
# round 2, fourth GPU pass: parity suite (radiative cases new), acceptance checks on the product (irradiation new), bench line, variants, ncu
TAG=${1:-r02_v4}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 900 python tests/checkers/run_reference_acceptance.py --gpu > gpurun_out/${TAG}_acceptance.log 2>&1; echo "acceptance rc=$?" >> gpurun_out/${TAG}_acceptance.log
tail -9 gpurun_out/${TAG}_acceptance.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-4500
STEPS=10 BENCH_ARGS="--no-tolerance-mode" bash tools/gpu_variants.sh ${TAG}_var $VARIANTS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 60 -c 12 -f -o gpurun_out/${TAG}_full_c5 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_full_c5.log 2>&1
ls -la gpurun_out/ | tail -14
