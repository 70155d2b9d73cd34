#!/bin/bash
# round 2, second 8-GPU pass (after the two-pass CFL and the collective monitor reductions): 8 ranks == 1 rank with the CFL in
# check mode (library, both halo modes, adiabatic; peer mode isothermal), C++ host --ranks 8, bench lines at 1 and 8 GPUs
TAG=${1:-r02_m8b}; N=8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
: > gpurun_out/${TAG}_check.log
export FARGO_B200_CFL=check
for cfg in "adiabatic_planet peer" "adiabatic_planet nccl" "isothermal_planet peer"; do
  set -- $cfg
  if [ $2 = nccl ]; then export FARGO_B200_HALO=nccl; else unset FARGO_B200_HALO; fi
  echo "== $1 $2, 8 ranks vs 1 (FARGO_B200_CFL=check)" >> gpurun_out/${TAG}_check.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/multi_gpu_check.py --physics $1 --nrad 512 --naz 512 --steps 12 2>&1 | grep '^{' | cut -c1-1200 >> gpurun_out/${TAG}_check.log
done
unset FARGO_B200_HALO FARGO_B200_CFL
cat gpurun_out/${TAG}_check.log | cut -c1-400
echo "== C++ host, --ranks 8 vs 1 (BASELINE configs[2] physics at 512 x 512)" > gpurun_out/${TAG}_host.log
timeout 600 python tools/diag_host_ranks.py baseline_config2_setup 8 2 Nrad=512 Naz=512 cps=-1 Nsnapshots=2 Nmonitor=2 WriteDiskQuantities=yes >> gpurun_out/${TAG}_host.log 2>&1
grep -c "differ        0" gpurun_out/${TAG}_host.log; grep "differ" gpurun_out/${TAG}_host.log | grep -v "differ        0" | head -5
for n in 1 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-tolerance-mode > gpurun_out/${TAG}_bench_n$n.log 2>&1
  fi
  grep '^{' gpurun_out/${TAG}_bench_n$n.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'GPUs', round(d['ms_per_step'],3), 'ms/step  checksum', d['checksum']['sha256'][:16], 'e2e', round(d['e2e']['value']/1e9,2), round(d['e2e']['at_one_snapshot_per_10K_steps']['value']/1e9,2), {k.split('<')[0].strip('('):round(v,3) for k,v in d['kernels_ms_per_step'].items() if v > 0.01})" || tail -5 gpurun_out/${TAG}_bench_n$n.log
done
