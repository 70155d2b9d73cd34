#!/bin/bash
# round 2, first GPU pass: parity suite, the reference's acceptance checks and BASELINE configs C2-C4 on the product,
# bench line, ncu --set full of one step at the bench size (C5)
TAG=${1:-r02_v1}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt 2>&1; nproc >> gpurun_out/${TAG}_smi.txt; free -g >> gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
timeout 900 python tests/checkers/run_reference_acceptance.py --gpu > gpurun_out/${TAG}_acceptance.log 2>&1; echo "acceptance rc=$?" >> gpurun_out/${TAG}_acceptance.log
tail -9 gpurun_out/${TAG}_acceptance.log
timeout 1500 python tests/checkers/baseline_configs_on_gpu.py ${TAG} > gpurun_out/${TAG}_baseline_configs.log 2>&1; echo "baseline configs rc=$?" >> gpurun_out/${TAG}_baseline_configs.log
tail -12 gpurun_out/${TAG}_baseline_configs.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-3000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 60 -c 14 -f -o gpurun_out/${TAG}_full_c5 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full_c5.log 2>&1
tail -2 gpurun_out/${TAG}_full_c5.log | cut -c1-300
ls -la gpurun_out/
