#!/bin/bash
# full GPU pass: parity tests, bench (with CPU baseline), ncu launch list, ncu --set full of one step
# usage: gpu_round.sh <tag>
TAG=${1:-round}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
# the reference's own acceptance checks (shock tube x4, spreading ring, cold disk + planet) on the product
timeout 900 python tests/checkers/run_reference_acceptance.py --gpu > gpurun_out/${TAG}_acceptance.log 2>&1; echo "acceptance rc=$?" >> gpurun_out/${TAG}_acceptance.log
tail -8 gpurun_out/${TAG}_acceptance.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-3500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s 40 -c 16 -f -o gpurun_out/${TAG}_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --nrad 4096 --naz 8192 > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-300
ls -la gpurun_out/
