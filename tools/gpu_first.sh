#!/bin/bash
# first GPU pass: parity tests, per-stage diagnostics, a short bench, the ncu launch list
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tests/checkers/gpu_diag.py > gpurun_out/diag.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c5.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_c5.log
tail -5 gpurun_out/bench_c5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
