#!/bin/bash
# parity tests + a short bench; usage: gpu_check.sh <tag> [bench args]
TAG=${1:-check}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
if ! grep -q "pytest rc=0" gpurun_out/${TAG}_pytest.log; then timeout 300 python tests/checkers/gpu_diag.py > gpurun_out/${TAG}_diag.log 2>&1; grep -v "ndiff=0 maxrel=0.00e+00 rings=\[\] | energy: ndiff=0 maxrel=0.00e+00 rings=\[\]$" gpurun_out/${TAG}_diag.log | head -40; fi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-3000
