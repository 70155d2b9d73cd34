#!/bin/bash
# time kernel-tuning variants of the library: gpu_variants.sh <tag> lib1.so lib2.so ...   (bench args via BENCH_ARGS)
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename $lib .so)
  FARGO_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/${TAG}_${name}.log 2>&1
  python - "$name" gpurun_out/${TAG}_${name}.log <<'PY'
import json,sys
name,path=sys.argv[1:3]
line=[l for l in open(path) if l.startswith('{')]
if not line: print(name,'FAILED',open(path).read()[-600:]); sys.exit()
d=json.loads(line[-1])
print(f"{name}: {d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']/1e9:.2f} G |", ' '.join(f"{k.split('<')[0].strip('(')}={v:.2f}" for k,v in d['kernels_ms_per_step'].items()))
PY
done
