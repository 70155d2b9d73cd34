#!/bin/bash
# round 2, last GPU pass: smoke(), the whole GPU suite, the full bench line (cpu_baseline, tolerance_mode)
TAG=${1:-r02_v23}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -2 gpurun_out/${TAG}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -2 gpurun_out/${TAG}_bench.log | cut -c1-6000
