#!/bin/bash
# round 2, fifth GPU pass: racecheck / memcheck of one small hydro run (the cp.async ring staging has no barrier by design),
# bench of the default library, variants of the viscosity kernel with its reciprocals hoisted to the top of the iteration
TAG=${1:-r02_v5}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/${TAG}_racecheck.log
tail -4 gpurun_out/${TAG}_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${TAG}_memcheck.log
tail -4 gpurun_out/${TAG}_memcheck.log
STEPS=10 BENCH_ARGS="--no-tolerance-mode" bash tools/gpu_variants.sh ${TAG}_var $VARIANTS
grep -h -o '"sha256": "[0-9a-f]*"' gpurun_out/${TAG}_var_*.log | sort | uniq -c
