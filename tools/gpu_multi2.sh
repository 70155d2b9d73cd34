#!/bin/bash
# 2-GPU pass: parity vs 1 GPU in both halo modes, then the bench at N GPUs in both modes. usage: gpu_multi2.sh <tag> <N>
TAG=$1; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
for mode in peer nccl; do
  if [ $mode = nccl ]; then export FARGO_B200_HALO=nccl; else unset FARGO_B200_HALO; fi
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}_$mode.log 2>&1
  echo "bench n=$N $mode rc=$?"; grep '^{' gpurun_out/${TAG}_bench_n${N}_$mode.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['ms_per_step'], d['config'].get('halo_exchange'), d['kernels_ms_per_step'])" || tail -20 gpurun_out/${TAG}_bench_n${N}_$mode.log
done
