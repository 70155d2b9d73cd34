#!/bin/bash
# compute-sanitizer memcheck + racecheck of the entry points added late in round 2
TAG=${1:-r02_v21}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_new_paths.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log
tail -5 gpurun_out/${TAG}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_new_paths.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck.log
tail -5 gpurun_out/${TAG}_racecheck.log
