#!/bin/bash
# round 2, ninth GPU pass: whole GPU suite (fargo_monitor_disk new), bench line, launch list, ncu --set full of one step at C5 size
TAG=${1:-r02_v13}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -2 gpurun_out/${TAG}_bench.log | cut -c1-5000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 78 -c 13 -f -o gpurun_out/${TAG}_full_c5 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_full_c5.log 2>&1
ls -la gpurun_out/ | tail -6
