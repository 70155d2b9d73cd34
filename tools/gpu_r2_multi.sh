#!/bin/bash
# round 2, N-GPU pass (N = number of GPUs of the box): multi-GPU parity tests (library: both halo modes; C++ host: --ranks N vs 1),
# bench lines at 1 and N GPUs with the N-independent checksum.   usage: gpu_r2_multi.sh <tag> <N>
TAG=${1:-r02_m2}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -rA > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_n1.log 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_n$N.log 2>&1
for n in 1 $N; do grep '^{' gpurun_out/${TAG}_bench_n$n.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'GPUs', round(d['ms_per_step'],3), 'ms/step  checksum', d['checksum']['sha256'][:16], 'e2e', round(d['e2e']['value']/1e9,2))"; done
