#!/bin/bash
# round 2, 8-GPU pass: N ranks == 1 rank (library, both halo modes, both physics; C++ host --ranks 8), strong-scaling bench lines
# at 1, 2, 4, 8 GPUs with the N-independent checksum.   usage: gpu_r2_multi8.sh <tag>
TAG=${1:-r02_m8}; N=8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
: > gpurun_out/${TAG}_check.log
for phys in isothermal_planet adiabatic_planet; do
  for halo in peer nccl; do
    if [ $halo = nccl ]; then export FARGO_B200_HALO=nccl; else unset FARGO_B200_HALO; fi
    echo "== $phys $halo, 8 ranks vs 1" >> gpurun_out/${TAG}_check.log
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tools/multi_gpu_check.py --physics $phys --nrad 512 --naz 512 --steps 12 2>&1 | grep '^{' | cut -c1-900 >> gpurun_out/${TAG}_check.log
  done
done
unset FARGO_B200_HALO
cat gpurun_out/${TAG}_check.log | cut -c1-300
echo "== C++ host, --ranks 8 vs 1 (BASELINE configs[2] physics at 512 x 512; accreting planet with DiskFeedback at 256 rings)" > gpurun_out/${TAG}_host.log
timeout 600 python tools/diag_host_ranks.py baseline_config2_setup 8 2 Nrad=512 Naz=512 cps=-1 Nsnapshots=2 Nmonitor=2 >> gpurun_out/${TAG}_host.log 2>&1
timeout 600 python tools/diag_host_ranks.py adia_accfb_20 8 3 Nrad=256 >> gpurun_out/${TAG}_host.log 2>&1
grep -c "differ        0" gpurun_out/${TAG}_host.log; grep "differ" gpurun_out/${TAG}_host.log | grep -v "differ        0" | head -5
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_n$n.log 2>&1
  fi
  grep '^{' gpurun_out/${TAG}_bench_n$n.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'GPUs', round(d['ms_per_step'],3), 'ms/step  checksum', d['checksum']['sha256'][:16], 'e2e', round(d['e2e']['value']/1e9,2), {k.split('<')[0].strip('('):round(v,3) for k,v in d['kernels_ms_per_step'].items() if v > 0.05})" || tail -5 gpurun_out/${TAG}_bench_n$n.log
done
