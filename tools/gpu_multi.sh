#!/bin/bash
# N-GPU pass: parity vs 1 GPU, then the bench at N GPUs.  usage: gpu_multi.sh <tag> <N>
TAG=$1; shift; N="$@"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for n in $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.log 2>&1
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$n.log 2>&1
  fi
  echo "bench n=$n rc=$?"; grep '^{' gpurun_out/${TAG}_bench_n$n.log | cut -c1-900 || tail -20 gpurun_out/${TAG}_bench_n$n.log
done
