#!/bin/bash
# N-GPU pass: parity of the peer-memory halo path vs 1 GPU, then the bench with cost-balanced and with equal cut points
TAG=$1; N=${2:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/multi_gpu_check.py --physics adiabatic_planet --nrad 512 --naz 512 --steps 12 > gpurun_out/${TAG}_check.log 2>&1
echo "check rc=$?"; grep '^{' gpurun_out/${TAG}_check.log | cut -c1-700
for split in balanced equal; do
  if [ $split = equal ]; then export FARGO_B200_SPLIT=equal; else unset FARGO_B200_SPLIT; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}_$split.log 2>&1
  echo "bench n=$N $split rc=$?"; grep '^{' gpurun_out/${TAG}_bench_n${N}_$split.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['ms_per_step'], d['value']/1e9, d['e2e']['value']/1e9, d['config'].get('halo_exchange')[:30], d['kernels_ms_per_step'])" || tail -20 gpurun_out/${TAG}_bench_n${N}_$split.log
done
